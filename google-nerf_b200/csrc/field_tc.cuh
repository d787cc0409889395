// field_tc.cuh -- device code of the fused field MLP forward (tcgen05 + TMEM) shared by field_tc.cu (the stand-alone
// forward / backward kernels) and render_tc.cu (whole rays in one kernel): weight-image offsets, SH-4, the ReLU / pack
// epilogues into canonical shared-memory tiles and the per-layer MMA issue.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

using namespace tc;
static_assert(ACT_LBO == 2064, "tile_load/tile_store hard-code the padded chunk stride");

// canonical weight image (halves): [W1 64xK1][W2 16x64][W3 64x32][W4 64x64][W5 16x64]
template <int K1>
struct Img {
    static constexpr int W1 = 0, W2 = 64 * K1, W3 = W2 + 1024, W4 = W3 + 2048, W5 = W4 + 4096, HALVES = W5 + 1024;
};

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

// ReLU + fp16 pack of 32 accumulator columns into chunks c0..c0+3 of this thread's row of a tile
__device__ __forceinline__ void relu32_to_tile(const float *v, unsigned char *tile, int r, int c0) {
    const __half2 zero2 = __float2half2_rn(0.f);
    #pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint4 o;
        __half2 *oh = reinterpret_cast<__half2 *>(&o);
        #pragma unroll
        for (int j = 0; j < 4; ++j) oh[j] = __hmax2(__floats2half2_rn(v[8 * c + 2 * j], v[8 * c + 2 * j + 1]), zero2);
        *reinterpret_cast<uint4 *>(tile + act_off(r, c0 + c)) = o;
    }
}
// ReLU of this thread's 64 accumulator columns -> its row of a [128 x 64] tile
__device__ __forceinline__ void relu_to_tile(const float *v, unsigned char *tile, int r) {
    relu32_to_tile(v, tile, r, 0);
    relu32_to_tile(v + 32, tile, r, 4);
}
// the same straight from TMEM in two 32-column halves (bounds the register footprint)
__device__ __forceinline__ void relu_tmem_to_tile(uint32_t tmem_work, unsigned char *tile, int r) {
    #pragma unroll
    for (int half32 = 0; half32 < 2; ++half32) {
        float v[32];
        tmem_ld32(tmem_work + half32 * 32, v);
        relu32_to_tile(v, tile, r, half32 * 4);
    }
}

// K-loop of one layer: D[128 x N] = A[128 x K] (K-major tile) * W[N x K]^T (K-major image); no commit
__device__ __forceinline__ void issue_layer_nc(uint32_t tmem_d, uint32_t a_addr, uint32_t w_addr, int N, int K) {
    const uint32_t idesc = make_idesc(128, N, 0, 0);
    const uint32_t w_lbo = (uint32_t)(N >> 3) * 128;
    for (int k = 0; k < K / 16; ++k) {
        const uint64_t da = make_desc(a_addr + (uint32_t)k * 2 * ACT_LBO, ACT_LBO, ACT_SBO);
        const uint64_t db = make_desc(w_addr + (uint32_t)k * 2 * w_lbo, w_lbo, 128);
        mma_f16_ss(tmem_d, da, db, idesc, k > 0 ? 1u : 0u);
    }
}
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t a_addr, uint32_t w_addr, int N, int K,
                                            uint64_t *bar) {
    issue_layer_nc(tmem_d, a_addr, w_addr, N, K);
    mma_commit(bar);
}

#define STEP_SYNC()            \
    do {                       \
        fence_async_smem();    \
        fence_before_sync();   \
        __syncthreads();       \
        fence_after_sync();    \
    } while (0)

template <int K1>
struct FieldFwSmem {
    __half w[Img<K1>::HALVES];                 // 20480 / 26624 B  canonical weight images
    unsigned char a0[(K1 / 8) * ACT_LBO];      //  8256 / 20640 B  encoded input tile  [128 x K1]
    unsigned char a1[TILE64_BYTES];            // 16512 B          hidden tile         [128 x 64]
    unsigned char a3[TILE32_BYTES];            //  8256 B          colour-net input    [128 x 32] = [SH16 | h16]
    uint64_t bar_w, bar_mma;
    uint32_t tmem_base;
};

