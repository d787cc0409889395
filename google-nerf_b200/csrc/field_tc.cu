// field_tc.cu -- the NGP field's dense part on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Replaces, for the HashGrid NGP hot path, the chain
//     xyz_encoder's FullyFusedMLP (32 -> 64 -> 16)  ->  TruncExp  ->  SH-4(dir)  ->  cat  ->  rgb_net (32 -> 64 -> 64 -> 3)
// of ngp_pl/models/networks.py:96-115 (five tcnn/torch launches plus casts in the reference) with ONE kernel
// forward and ONE kernel backward.  Layer activations never leave the SM between layers: each layer is one
// tcgen05.mma group (M = 128 samples, N = layer width, K = 16 per instruction) whose fp32 accumulator lives in
// TMEM; the epilogue threads pull their row with tcgen05.ld, apply ReLU / exp / sigmoid, and write the fp16
// row straight back into the shared-memory operand tile of the next layer.  Weights (20 KB fp16, laid out
// once in the UMMA canonical layout by field_pack_weights) arrive with one TMA bulk copy per CTA.
//
// Work decomposition: CTA = 128 threads = 128 samples (thread r <-> sample row r <-> TMEM lane r); CTAs are
// persistent over tiles; several CTAs per SM overlap one CTA's epilogue with another's MMA.  Activations that
// the backward pass needs are copied out of the operand tiles with fully coalesced 16-byte stores while the next
// layer's MMA is in flight.  The arithmetic (20.5 kFLOP/sample) is far below the tensor roofline; the kernels are
// bound by the activation bytes saved / re-read for the backward pass (DESIGN.md "Field MLP").
#include "common.cuh"
#include "tc_common.cuh"

using namespace tc;
static_assert(ACT_LBO == 2064, "tile_load/tile_store hard-code the padded chunk stride");

// canonical weight image (halves): [W1 64x32][W2 16x64][W3 64x32][W4 64x64][W5 16x64]
#define IMG_W1 0
#define IMG_W2 2048
#define IMG_W3 3072
#define IMG_W4 5120
#define IMG_W5 9216
#define IMG_HALVES 10240

__global__ void __launch_bounds__(256) field_pack_weights_kernel(const __half *__restrict__ sigma_w,
                                                                 const __half *__restrict__ rgb_w,
                                                                 __half *__restrict__ image) {
    // flat row-major (out,in) matrices -> canonical K-major no-swizzle images
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < IMG_HALVES; e += gridDim.x * blockDim.x) {
        const __half *src; int base, N, K, idx;
        if (e < 2048)      { src = sigma_w;        base = IMG_W1; N = 64; K = 32; idx = e; }
        else if (e < 3072) { src = sigma_w + 2048; base = IMG_W2; N = 16; K = 64; idx = e - 2048; }
        else if (e < 5120) { src = rgb_w;          base = IMG_W3; N = 64; K = 32; idx = e - 3072; }
        else if (e < 9216) { src = rgb_w + 2048;   base = IMG_W4; N = 64; K = 64; idx = e - 5120; }
        else               { src = rgb_w + 6144;   base = IMG_W5; N = 16; K = 64; idx = e - 9216; }
        const int n = idx / K, k = idx - n * K;
        image[base + w_off_halves(n, k, N)] = src[idx];
    }
}

extern "C" int b2n_field_pack_weights(const b2n_half *sigma_w, const b2n_half *rgb_w, b2n_half *image, void *stream) {
    field_pack_weights_kernel<<<10, 256, 0, (cudaStream_t)stream>>>((const __half *)sigma_w, (const __half *)rgb_w,
                                                                    (__half *)image);
    B2N_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sh4_eval_dev(float x, float y, float z, float *o) {
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    o[0] = 0.28209479177387814f;
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    o[10] = 2.8906114426405538f * xy * z;
    o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    o[14] = 1.4453057213202769f * z * (x2 - y2);
    o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

__device__ __forceinline__ uint4 pack8(const float *v) {
    uint4 p;
    __half2 *h = reinterpret_cast<__half2 *>(&p);
    h[0] = __floats2half2_rn(v[0], v[1]); h[1] = __floats2half2_rn(v[2], v[3]);
    h[2] = __floats2half2_rn(v[4], v[5]); h[3] = __floats2half2_rn(v[6], v[7]);
    return p;
}

// SH-4 of the normalised direction of this thread's row -> chunks 0,1 of a [128 x 32] tile
__device__ __forceinline__ void sh_to_tile(const float *__restrict__ dirs, int64_t row, bool live, unsigned char *tile,
                                           int r) {
    float dx = 0.f, dy = 0.f, dz = 1.f;
    if (live) { dx = __ldg(dirs + 3 * row); dy = __ldg(dirs + 3 * row + 1); dz = __ldg(dirs + 3 * row + 2); }
    const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
    float sh[16];
    sh4_eval_dev(dx * inv, dy * inv, dz * inv, sh);
    if (!live) {
        #pragma unroll
        for (int i = 0; i < 16; ++i) sh[i] = 0.f;
    }
    *reinterpret_cast<uint4 *>(tile + act_off(r, 0)) = pack8(sh);
    *reinterpret_cast<uint4 *>(tile + act_off(r, 1)) = pack8(sh + 8);
}

// ReLU + fp16 pack of this thread's 64 accumulator columns into its row of a [128 x 64] tile
__device__ __forceinline__ void relu_to_tile(const float *v, unsigned char *tile, int r) {
    const __half2 zero2 = __float2half2_rn(0.f);
    #pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint4 o;
        __half2 *oh = reinterpret_cast<__half2 *>(&o);
        #pragma unroll
        for (int j = 0; j < 4; ++j) oh[j] = __hmax2(__floats2half2_rn(v[8 * c + 2 * j], v[8 * c + 2 * j + 1]), zero2);
        *reinterpret_cast<uint4 *>(tile + act_off(r, c)) = o;
    }
}

// issue the K-loop of one layer: D[128 x N] = A[128 x K] (K-major tile) * W[N x K]^T (K-major image)
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t a_addr, uint32_t w_addr, int N, int K,
                                            uint64_t *bar) {
    const uint32_t idesc = make_idesc(128, N, 0, 0);
    const uint32_t w_lbo = (uint32_t)(N >> 3) * 128;
    for (int k = 0; k < K / 16; ++k) {
        const uint64_t da = make_desc(a_addr + (uint32_t)k * 2 * ACT_LBO, ACT_LBO, ACT_SBO);
        const uint64_t db = make_desc(w_addr + (uint32_t)k * 2 * w_lbo, w_lbo, 128);
        mma_f16_ss(tmem_d, da, db, idesc, k > 0 ? 1u : 0u);
    }
    mma_commit(bar);
}

#define STEP_SYNC()            \
    do {                       \
        fence_async_smem();    \
        fence_before_sync();   \
        __syncthreads();       \
        fence_after_sync();    \
    } while (0)

struct FieldFwSmem {
    __half w[IMG_HALVES];                      // 20480 B  canonical weight images
    unsigned char a0[TILE32_BYTES];            //  8256 B  encoded input tile  [128 x 32]
    unsigned char a1[TILE64_BYTES];            // 16512 B  hidden tile         [128 x 64]
    unsigned char a3[TILE32_BYTES];            //  8256 B  colour-net input    [128 x 32] = [SH16 | h16]
    uint64_t bar_w, bar_mma;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(128, 4) field_mlp_fw_kernel(const __half *__restrict__ enc,
                                                              const float *__restrict__ dirs,
                                                              const __half *__restrict__ image, int64_t n,
                                                              const int32_t *__restrict__ n_dev,
                                                              float *__restrict__ sigmas, float *__restrict__ rgbs,
                                                              __half *__restrict__ hid_s, __half *__restrict__ h_out,
                                                              __half *__restrict__ hid_r) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    FieldFwSmem &S = *reinterpret_cast<FieldFwSmem *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int64_t n_alloc = n;
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
        mbar_expect_tx(&S.bar_w, IMG_HALVES * 2);
        bulk_g2s(S.w, image, IMG_HALVES * 2, &S.bar_w);
    }
    mbar_wait(&S.bar_w, 0);

    const uint32_t w_addr = smem_u32(S.w), a0 = smem_u32(S.a0), a1 = smem_u32(S.a1), a3 = smem_u32(S.a3);
    const bool density_only = (rgbs == nullptr);
    uint32_t phase = 0;

    // the encoded tile of the NEXT tile is fetched with cp.async into a0 as soon as layer 1 has consumed the
    // current one (four layers ahead of its use); the first tile is fetched here
    tile_load_async<4>(S.a0, enc + (int64_t)blockIdx.x * 128 * 32, n - (int64_t)blockIdx.x * 128, tid);
    cp_async_commit();
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * 128, row = row0 + tid, rows_valid = n - row0;
        const bool live = row < n;
        const int64_t next = tile + gridDim.x;
        // ---- stage 0: SH(dir) into the colour-net operand tile; the encoded features are already in flight
        if (!density_only) sh_to_tile(dirs, row, live, S.a3, tid);
        cp_async_wait_all();
        STEP_SYNC();
        // ---- layer 1: enc(32) -> 64, ReLU
        if (tid == 0) issue_layer(tmem, a0, w_addr + IMG_W1 * 2, 64, 32, &S.bar_mma);
        mbar_wait(&S.bar_mma, phase); phase ^= 1;
        fence_after_sync();
        if (next < n_tiles) {                         // a0 is free again: next tile's encoded features
            tile_load_async<4>(S.a0, enc + next * 128 * 32, n - next * 128, tid);
            cp_async_commit();
        }
        {
            float v[64];
            tmem_ld64(tmem_row, v);
            relu_to_tile(v, S.a1, tid);
        }
        STEP_SYNC();
        // ---- layer 2: 64 -> 16 (h); sigma = exp(h[0])   (TruncExp forward, custom_functions.py:165-167)
        if (tid == 0) issue_layer(tmem, a1, w_addr + IMG_W2 * 2, 16, 64, &S.bar_mma);
        if (hid_s != nullptr) tile_store<8>(S.a1, hid_s + row0 * 64, rows_valid, tid);   // overlaps the MMA
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
                sigmas[row] = expf(h0);
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
        if (tid == 0) issue_layer(tmem, a3, w_addr + IMG_W3 * 2, 64, 32, &S.bar_mma);
        mbar_wait(&S.bar_mma, phase); phase ^= 1;
        fence_after_sync();
        {
            float v[64];
            tmem_ld64(tmem_row, v);
            relu_to_tile(v, S.a1, tid);
        }
        STEP_SYNC();
        // ---- layer 4: 64 -> 64, ReLU
        if (tid == 0) issue_layer(tmem, a1, w_addr + IMG_W4 * 2, 64, 64, &S.bar_mma);
        if (hid_r != nullptr) tile_store<8>(S.a1, hid_r + row0 * 64, rows_valid, tid);
        mbar_wait(&S.bar_mma, phase); phase ^= 1;
        __syncthreads();       // every thread's copy-out of a1 is done before anyone overwrites its row below
        fence_after_sync();
        {
            float v[64];
            tmem_ld64(tmem_row, v);
            relu_to_tile(v, S.a1, tid);
        }
        STEP_SYNC();
        // ---- layer 5: 64 -> 3 (padded 16), sigmoid
        if (tid == 0) issue_layer(tmem, a1, w_addr + IMG_W5 * 2, 16, 64, &S.bar_mma);
        if (hid_r != nullptr) tile_store<8>(S.a1, hid_r + (n_alloc + row0) * 64, rows_valid, tid);
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

extern "C" int b2n_field_mlp_fw(const b2n_half *enc, const float *dirs, const b2n_half *image, int64_t n,
                                const int32_t *n_dev, float *sigmas, float *rgbs, b2n_half *hid_s, b2n_half *h,
                                b2n_half *hid_r, void *stream) {
    B2N_CHECK_ARG(((uintptr_t)enc & 15) == 0 && ((uintptr_t)image & 15) == 0, "enc / image must be 16-byte aligned");
    if (n <= 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(field_mlp_fw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FieldFwSmem) + 256);
        attr_set = true;
    }
    field_mlp_fw_kernel<<<b2n_grid((n + 127) / 128, 4), 128, sizeof(FieldFwSmem) + 256, (cudaStream_t)stream>>>(
        (const __half *)enc, dirs, (const __half *)image, n, n_dev, sigmas, rgbs, (__half *)hid_s, (__half *)h,
        (__half *)hid_r);
    B2N_LAUNCH_CHECK();
    return 0;
}

// ================================================================================================
// Backward.  Per 128-sample tile, five steps (layer 5 .. layer 1); each step issues, back to back,
//   wgrad:  dW += g^T . act       (M = 64, K = 128 samples; both operands are the SAME shared-memory tiles the
//                                  forward/dgrad MMAs use, read MN-major; accumulators persist in TMEM across
//                                  all tiles of the CTA and are flushed once with red.global.add)
//   dgrad:  g_prev = g . W        (M = 128 samples; W read MN-major from the canonical weight image)
// then the epilogue threads apply ReLU' / TruncExp' and write the fp16 gradient tile of the next step in place.
// Activation tiles of the next step are fetched from global memory (coalesced) while the current MMAs run.
//
// Latency hiding: the chain of a tile is ~10 dependent MMA groups + global loads, and the 160 TMEM columns of
// weight-gradient accumulators cap co-resident CTAs, so ONE CTA per SM runs BW_GROUPS = 4 independent 128-thread
// groups, each with its own tile, shared-memory tiles, mbarrier and 64-column dgrad accumulator, all accumulating
// into the SAME weight-gradient columns.  MMA issue is serialised by a shared-memory lock bracketed with
// tcgen05 fences so that accumulation into the shared columns is ordered.
// TMEM columns: [64 g, 64 g + 64) dgrad of group g | 256.. dW5^T(16) dW4(64) dW3(32) dW2^T(16) dW1(32).
#define BW_GROUPS 4
// Phase trace (debug builds only, -DB2N_BW_TRACE): clock64() of CTA 0 / group 0 / thread 64 at every phase boundary.
#ifdef B2N_BW_TRACE
__device__ long long g_bw_trace[64 * 16];
#define TRACE(slot) do { if (blockIdx.x == 0 && threadIdx.x == 64 && trace_it < 64) g_bw_trace[trace_it * 16 + (slot)] = clock64(); } while (0)
extern "C" __attribute__((visibility("default"))) int b2n_debug_bw_trace(long long *out) { return (int)cudaMemcpyFromSymbol(out, g_bw_trace, sizeof(g_bw_trace)); }
#else
#define TRACE(slot) do { } while (0)
#endif
struct FieldBwSmem {
    __half w[IMG_HALVES];                               //  20480 B
    unsigned char act[BW_GROUPS][2][TILE64_BYTES];      // 132096 B  activation tiles (ping-pong per group)
    unsigned char g[BW_GROUPS][TILE64_BYTES];           //  66048 B  gradient tile (in place per group)
    int32_t idx[BW_GROUPS][2][128];                     //   4096 B  sample rows of the current / next tile
    uint64_t bar_w, bar_mma[BW_GROUPS];
    uint32_t tmem_base;
    int lock;
};

#define TM_WG 256
#define TM_DW5T (TM_WG + 0)
#define TM_DW4 (TM_WG + 16)
#define TM_DW3 (TM_WG + 80)
#define TM_DW2T (TM_WG + 112)
#define TM_DW1 (TM_WG + 128)
#define TM_WG_COLS 160

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
// dgrad: D[128 x N_in] = G_tile[128 x K_out] (K-major) * W[K_out x N_in] (W image has K_out rows: MN-major B)
__device__ __forceinline__ void issue_dgrad(uint32_t tmem_d, uint32_t g_tile, uint32_t w_img, int N_in, int K_out) {
    const uint32_t idesc = make_idesc(128, N_in, 0, 1);
    const uint32_t w_lbo = (uint32_t)(K_out >> 3) * 128;          // byte stride between in-chunks of the image
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

// MMA issue from the four group leaders is NOT serialised: tcgen05.mma executes in the tensor pipe one instruction
// at a time and D += A*B is a single read-modify-write of TMEM inside that pipe, so accumulations into the shared
// weight-gradient columns from different issuing threads commute (only their order, i.e. fp32 summation order, is
// unspecified).  B2N_BW_LOCK=1 at compile time restores a shared-memory lock around every issue sequence.
#ifndef B2N_BW_LOCK
#define B2N_BW_LOCK 0
#endif
__device__ __forceinline__ void issue_lock(int *lock) {
#if B2N_BW_LOCK
    while (atomicCAS(lock, 0, 1) != 0) __nanosleep(32);
#endif
    fence_after_sync();
}
__device__ __forceinline__ void issue_unlock(int *lock) {
    fence_before_sync();
#if B2N_BW_LOCK
    __threadfence_block();
    atomicExch(lock, 0);
#endif
}

// g_next = (act > 0) ? dgrad : 0 for this thread's 64 columns (read from TMEM in two halves to bound registers).
// The mask is applied in the packed fp16 domain: cvt.rn.f16x2 of the gradient pair, HSETP2-style (act > 0) -> {1,0}
// and one HMUL2 -- three instructions per two values instead of convert / compare / select / pack per value (the
// epilogue is issue bound when several groups reach it together).
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

__global__ void __launch_bounds__(128 * BW_GROUPS, 1) field_mlp_bw_kernel(
    const float *__restrict__ dL_dsigmas, const float *__restrict__ dL_drgbs, const __half *__restrict__ enc,
    const float *__restrict__ dirs, const __half *__restrict__ image, int64_t n, const int32_t *__restrict__ n_dev,
    const float *__restrict__ rgbs, const __half *__restrict__ hid_s, const __half *__restrict__ h_in,
    const __half *__restrict__ hid_r, float grad_scale, __half *__restrict__ dL_denc,
    float *__restrict__ grad_sigma_w, float *__restrict__ grad_rgb_w, const int32_t *__restrict__ sample_idx,
    int64_t n_alloc) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    FieldBwSmem &S = *reinterpret_cast<FieldBwSmem *>(smem_raw);
    const int grp = threadIdx.x >> 7, tid = threadIdx.x & 127, warp = tid >> 5, lane = tid & 31;
    n = b2n_eff_n(n, n_dev);
    const int64_t n_tiles = (n + 127) / 128;
    if ((int64_t)blockIdx.x * BW_GROUPS >= n_tiles) return;          // uniform per CTA

    if (threadIdx.x == 0) {
        mbar_init(&S.bar_w, 1);
        for (int gq = 0; gq < BW_GROUPS; ++gq) mbar_init(&S.bar_mma[gq], 1);
        S.lock = 0;
        mbar_init_fence();
    }
    if (threadIdx.x < 32) tmem_alloc<512>(&S.tmem_base);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = S.tmem_base;
    const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);      // warp % 4 selects the TMEM lane quadrant
    const uint32_t tmem_work = tmem_row + grp * 64;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&S.bar_w, IMG_HALVES * 2);
        bulk_g2s(S.w, image, IMG_HALVES * 2, &S.bar_w);
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
    const uint32_t act_a[2] = {smem_u32(S.act[grp][0]), smem_u32(S.act[grp][1])};
    const uint32_t g_a = smem_u32(S.g[grp]);
    unsigned char *ACT[2] = {S.act[grp][0], S.act[grp][1]};
    unsigned char *G = S.g[grp];
    uint64_t *bar = &S.bar_mma[grp];
    const uint32_t tm_d = tmem + grp * 64;                               // dgrad accumulator (lane 0 base) of this group
    uint32_t phase = 0;
    // The four groups start together and would march through their (identical) phases in lock step -- all issuing
    // MMAs, then all in the issue-bound epilogue.  A one-off skew per group interleaves them (0 / 300 /
    // 1000 / 2400 ns measured within 2 us of each other).
#ifndef BW_SKEW_NS
#define BW_SKEW_NS 1000
#endif
    if (grp) __nanosleep(grp * BW_SKEW_NS);

    // Software pipeline: every global->shared tile copy is a cp.async issued ONE STEP AHEAD of its consumer (the
    // buffer roles swap with the tile parity q so that the next tile's first two tiles can be fetched during
    // step E), and the per-row scalars travel in registers.  Buffers of tile parity q: X = ACT[q], Y = ACT[1-q]:
    //   step A: X = hid_r2   step B: Y = hid_r1   step C: X = [SH|h]   step D: Y = hid_s   step E: X = enc
    const int64_t tile_stride = (int64_t)gridDim.x * BW_GROUPS;
    int64_t tile = (int64_t)blockIdx.x * BW_GROUPS + grp;
    int q = 0;
    float c5[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};          // rgb (3) and dL/drgb (3) of this thread's row
    // Row r of a tile is sample IDX[r]: the identity, or an entry of the compacted alive list (sample_idx).
    int32_t *IDX[2] = {S.idx[grp][0], S.idx[grp][1]};
    int ib = 0;
    if (tile < n_tiles) {
        const int64_t row0 = tile * 128, row = row0 + tid, rows_valid = n - row0;
        const int32_t my_s = (row < n) ? (sample_idx ? __ldg(sample_idx + row) : (int32_t)row) : 0;
        IDX[0][tid] = my_s;
        group_sync(grp);
        tile_gather_async<8>(ACT[0], hid_r + n_alloc * 64, IDX[0], rows_valid, tid);
        tile_gather_async<8>(ACT[1], hid_r, IDX[0], rows_valid, tid);
        cp_async_commit();
        if (row < n) {
            #pragma unroll
            for (int c = 0; c < 3; ++c) { c5[c] = __ldg(rgbs + 3 * my_s + c); c5[3 + c] = __ldg(dL_drgbs + 3 * my_s + c); }
        }
    }
    int trace_it = -1;
    for (; tile < n_tiles; tile += tile_stride, q ^= 1, ib ^= 1) {
        ++trace_it; (void)trace_it;
        TRACE(0);
        const int64_t row0 = tile * 128, row = row0 + tid, rows_valid = n - row0;
        const bool live = row < n;
        unsigned char *X = ACT[q], *Y = ACT[q ^ 1];
        const uint32_t x_a = act_a[q], y_a = act_a[q ^ 1];
        const int64_t next = tile + tile_stride;
        const bool has_next = next < n_tiles;
        const int64_t nrow0 = next * 128, nrow = nrow0 + tid, nrows_valid = n - nrow0;
        const int64_t my_s = IDX[ib][tid];                                     // this thread's sample row
        // ---- step A prologue: g5 = dL_drgb * sigmoid'(rgb) -> G [128x16]
        float dsx = 0.f, dsy = 0.f, dsz = 1.f, dsig = 0.f;
        {
            float g5[16];
            #pragma unroll
            for (int i = 0; i < 16; ++i) g5[i] = 0.f;
            #pragma unroll
            for (int c = 0; c < 3; ++c) g5[c] = live ? c5[3 + c] * c5[c] * (1.0f - c5[c]) : 0.f;
            *reinterpret_cast<uint4 *>(G + act_off(tid, 0)) = pack8(g5);
            *reinterpret_cast<uint4 *>(G + act_off(tid, 1)) = pack8(g5 + 8);
            if (live) {       // needed in steps B / C: in flight while step A runs
                dsx = __ldg(dirs + 3 * my_s); dsy = __ldg(dirs + 3 * my_s + 1); dsz = __ldg(dirs + 3 * my_s + 2);
                dsig = __ldg(dL_dsigmas + my_s);
            }
            cp_async_wait_all();                                              // hid_r2 (X) and hid_r1 (Y) have landed
        }
        TRACE(1);
        GROUP_STEP_SYNC();
        TRACE(2);
        // ---- step A: layer 5 (64 -> 16)
        if (tid == 0) {
            issue_lock(&S.lock);
            issue_wgrad(tmem + TM_DW5T, x_a, g_a, 16);                        // dW5^T[in][out] += hid_r2^T . g5
            issue_dgrad(tm_d, g_a, w_addr + IMG_W5 * 2, 64, 16);             // g4 = g5 . W5
            mma_commit(bar);
            issue_unlock(&S.lock);
        }
        mbar_wait(bar, phase); phase ^= 1;
        fence_after_sync();
        TRACE(3);
        relu_bw_epilogue(tmem_work, X, G, tid);                              // g4 = . * (hid_r2 > 0)
        TRACE(4);
        GROUP_STEP_SYNC();
        TRACE(5);
        // ---- step B: layer 4 (64 -> 64)
        if (tid == 0) {
            issue_lock(&S.lock);
            issue_wgrad(tmem + TM_DW4, g_a, y_a, 64);                         // dW4[out][in] += g4^T . hid_r1
            issue_dgrad(tm_d, g_a, w_addr + IMG_W4 * 2, 64, 64);              // g3 = g4 . W4
            mma_commit(bar);
            issue_unlock(&S.lock);
        }
        {   // X is free (step A done): colour-net input [SH16 | h16] for step C
            cp_async16(X + act_off(tid, 2), h_in + my_s * 16, live);
            cp_async16(X + act_off(tid, 3), h_in + my_s * 16 + 8, live);
            cp_async_commit();
            const float inv = 1.0f / sqrtf(dsx * dsx + dsy * dsy + dsz * dsz);
            float sh[16];
            sh4_eval_dev(dsx * inv, dsy * inv, dsz * inv, sh);
            if (!live) {
                #pragma unroll
                for (int i = 0; i < 16; ++i) sh[i] = 0.f;
            }
            *reinterpret_cast<uint4 *>(X + act_off(tid, 0)) = pack8(sh);
            *reinterpret_cast<uint4 *>(X + act_off(tid, 1)) = pack8(sh + 8);
        }
        TRACE(6);
        mbar_wait(bar, phase); phase ^= 1;
        fence_after_sync();
        TRACE(7);
        relu_bw_epilogue(tmem_work, Y, G, tid);                              // g3 = . * (hid_r1 > 0)
        TRACE(8);
        cp_async_wait_all();
        TRACE(9);
        GROUP_STEP_SYNC();
        TRACE(10);
        // ---- step C: layer 3 (32 -> 64)
        if (tid == 0) {
            issue_lock(&S.lock);
            issue_wgrad(tmem + TM_DW3, g_a, x_a, 32);                         // dW3[out][in] += g3^T . [SH|h]
            issue_dgrad(tm_d, g_a, w_addr + IMG_W3 * 2, 32, 64);              // g_in3 = g3 . W3
            mma_commit(bar);
            issue_unlock(&S.lock);
        }
        tile_gather_async<8>(Y, hid_s, IDX[ib], rows_valid, tid);             // Y is free (step B done): hid_s for step D
        cp_async_commit();
        mbar_wait(bar, phase); phase ^= 1;
        fence_after_sync();
        TRACE(12);
        {
            float v[16];
            tmem_ld16(tmem_work + 16, v);                                    // columns 16..31 = dL/dh from the colour net
            // TruncExp backward (custom_functions.py:171-173) joins on channel 0
            const __half2 hh = *reinterpret_cast<const __half2 *>(X + act_off(tid, 2));
            const float h0 = __low2float(hh);
            if (live) v[0] += dsig * expf(fminf(fmaxf(h0, -15.f), 15.f));
            *reinterpret_cast<uint4 *>(G + act_off(tid, 0)) = pack8(v);
            *reinterpret_cast<uint4 *>(G + act_off(tid, 1)) = pack8(v + 8);
        }
        cp_async_wait_all();
        GROUP_STEP_SYNC();
        TRACE(13);
        // ---- step D: layer 2 (64 -> 16)
        if (tid == 0) {
            issue_lock(&S.lock);
            issue_wgrad(tmem + TM_DW2T, y_a, g_a, 16);                        // dW2^T[in][out] += hid_s^T . g2
            issue_dgrad(tm_d, g_a, w_addr + IMG_W2 * 2, 64, 16);              // g1 = g2 . W2
            mma_commit(bar);
            issue_unlock(&S.lock);
        }
        tile_gather_async<4>(X, enc, IDX[ib], rows_valid, tid);               // X is free (step C done): enc for step E
        if (has_next)                                                         // next tile's sample rows (visible after the sync below)
            IDX[ib ^ 1][tid] = (nrow < n) ? (sample_idx ? __ldg(sample_idx + nrow) : (int32_t)nrow) : 0;
        cp_async_commit();
        mbar_wait(bar, phase); phase ^= 1;
        fence_after_sync();
        TRACE(14);
        relu_bw_epilogue(tmem_work, Y, G, tid);                              // g1 = . * (hid_s > 0)
        cp_async_wait_all();
        GROUP_STEP_SYNC();
        TRACE(15);
        // ---- step E: layer 1 (32 -> 64)
        if (tid == 0) {
            issue_lock(&S.lock);
            issue_wgrad(tmem + TM_DW1, g_a, x_a, 32);                         // dW1[out][in] += g1^T . enc
            issue_dgrad(tm_d, g_a, w_addr + IMG_W1 * 2, 32, 64);              // g_enc = g1 . W1
            mma_commit(bar);
            issue_unlock(&S.lock);
        }
        if (has_next) {       // Y is free (step D done): the next tile's step-A tile and its per-row scalars
            tile_gather_async<8>(Y, hid_r + n_alloc * 64, IDX[ib ^ 1], nrows_valid, tid);
            cp_async_commit();
            #pragma unroll
            for (int c = 0; c < 6; ++c) c5[c] = 0.f;
            if (nrow < n) {
                const int64_t ns = IDX[ib ^ 1][tid];
                #pragma unroll
                for (int c = 0; c < 3; ++c) { c5[c] = __ldg(rgbs + 3 * ns + c); c5[3 + c] = __ldg(dL_drgbs + 3 * ns + c); }
            }
        }
        mbar_wait(bar, phase); phase ^= 1;
        fence_after_sync();
        if (has_next) {       // X is free (step E done): the next tile's step-B tile
            tile_gather_async<8>(X, hid_r, IDX[ib ^ 1], nrows_valid, tid);
            cp_async_commit();
        }
        {   // dL/denc: own row -> G tile (its readers, the step-E MMAs, are done), then one coalesced copy to global
            float v[32];
            tmem_ld32(tmem_work, v);
            #pragma unroll
            for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4 *>(G + act_off(tid, c)) = pack8(v + 8 * c);
        }
        fence_before_sync();
        group_sync(grp);
        fence_after_sync();
        tile_store<4>(G, dL_denc + row0 * 32, rows_valid, tid);
        group_sync(grp);                 // the next tile's prologue overwrites G rows other threads just copied out
        TRACE(11);
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
                    if (col < TM_DW4)       atomicAdd(grad_rgb_w + 6144 + (col - TM_DW5T) * 64 + m, val);   // dW5^T[in=m][out]
                    else if (col < TM_DW3)  atomicAdd(grad_rgb_w + 2048 + m * 64 + (col - TM_DW4), val);    // dW4[out=m][in]
                    else if (col < TM_DW2T) atomicAdd(grad_rgb_w + m * 32 + (col - TM_DW3), val);           // dW3[out=m][in]
                    else if (col < TM_DW1)  atomicAdd(grad_sigma_w + 2048 + (col - TM_DW2T) * 64 + m, val); // dW2^T[in=m][out]
                    else                    atomicAdd(grad_sigma_w + m * 32 + (col - TM_DW1), val);         // dW1[out=m][in]
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

extern "C" int b2n_field_mlp_bw(const float *dL_dsigmas, const float *dL_drgbs, const b2n_half *enc, const float *dirs,
                                const b2n_half *image, int64_t n, const int32_t *n_dev, const float *rgbs,
                                const b2n_half *hid_s, const b2n_half *h, const b2n_half *hid_r, float grad_scale,
                                b2n_half *dL_denc, float *grad_sigma_w, float *grad_rgb_w, const int32_t *sample_idx,
                                int64_t n_alloc, void *stream) {
    B2N_CHECK_ARG(hid_s && h && hid_r && rgbs && dL_denc && grad_sigma_w && grad_rgb_w, "saved activations and outputs are required");
    if (n <= 0) return 0;
    if (sample_idx == nullptr || n_alloc <= 0) n_alloc = n;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(field_mlp_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FieldBwSmem) + 256);
        attr_set = true;
    }
    const int64_t n_tiles = (n + 127) / 128;
    field_mlp_bw_kernel<<<b2n_grid((n_tiles + BW_GROUPS - 1) / BW_GROUPS, 1), 128 * BW_GROUPS, sizeof(FieldBwSmem) + 256,
                          (cudaStream_t)stream>>>(
        dL_dsigmas, dL_drgbs, (const __half *)enc, dirs, (const __half *)image, n, n_dev, rgbs, (const __half *)hid_s,
        (const __half *)h, (const __half *)hid_r, grad_scale, (__half *)dL_denc, grad_sigma_w, grad_rgb_w, sample_idx,
        n_alloc);
    B2N_LAUNCH_CHECK();
    return 0;
}
