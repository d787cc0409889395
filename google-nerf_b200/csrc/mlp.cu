// mlp.cu -- 64-wide fully fused MLP forward / backward (FullyFusedMLP replacement), CUDA-core version.
// Replaces tcnn.Network / the network half of tcnn.NetworkWithInputEncoding
// (ngp_pl/models/networks.py:54-60,72-83): ReLU hidden layers, no biases, fp16 weights and activations,
// fp32 accumulation, output padded to 16 columns, optional sigmoid.
//
// One CTA = 128 samples, one thread = one sample row (activations stay in the thread's shared-memory row
// between layers, so no inter-layer synchronisation); weights live in shared memory and are read as
// warp-wide broadcasts.  The backward kernel is persistent: weight-gradient partial sums live in fp32
// shared memory across tiles and are flushed with one red.global.add per weight per CTA.
// The tensor-core (tcgen05) fused field kernel in field_tc.cu supersedes this on the NGP hot path; this
// file remains the generic path of the `tinycudann` shim (arbitrary in_width, standalone Network).
#include "common.cuh"

#define TILE 128
#define HID 64
#define OUTW 16
#define MAX_IN 80
#define ROWP 8  // row padding (halves) of the activation tile: keeps 16-byte alignment, spreads banks

struct MlpShape {
    int in_width, n_hidden, out_act;
    int w_off[4];  // element offsets of each layer's matrix in the flat weights
    int n_layers;  // n_hidden + 1
    int n_weights;
};

static int make_shape(int in_width, int n_hidden, int out_act, MlpShape &s) {
    B2N_CHECK_ARG(in_width % 16 == 0 && in_width >= 16 && in_width <= MAX_IN, "in_width must be 16..80, multiple of 16");
    B2N_CHECK_ARG(n_hidden >= 1 && n_hidden <= 3, "n_hidden must be 1..3");
    B2N_CHECK_ARG(out_act == 0 || out_act == 1, "output_activation must be 0 (none) or 1 (sigmoid)");
    s.in_width = in_width; s.n_hidden = n_hidden; s.out_act = out_act; s.n_layers = n_hidden + 1;
    int off = 0;
    for (int l = 0; l < s.n_layers; ++l) {
        s.w_off[l] = off;
        const int o = (l == n_hidden) ? OUTW : HID, i = (l == 0) ? in_width : HID;
        off += o * i;
    }
    s.n_weights = off;
    return 0;
}

__device__ __forceinline__ int layer_in(const MlpShape &s, int l) { return l == 0 ? s.in_width : HID; }
__device__ __forceinline__ int layer_out(const MlpShape &s, int l) { return l == s.n_hidden ? OUTW : HID; }

// acc[o] += sum_k act[k] * Wt[k][o]   (Wt = transposed weights in smem, row k holds NOUT halves)
template <int NOUT>
__device__ __forceinline__ void row_times_wt(const __half *act_row, int n_in, const __half *wt, float *acc) {
    for (int k0 = 0; k0 < n_in; k0 += 8) {
        const uint4 av = *reinterpret_cast<const uint4 *>(act_row + k0);
        const __half2 *ah = reinterpret_cast<const __half2 *>(&av);
        #pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const float a = (kk & 1) ? __high2float(ah[kk >> 1]) : __low2float(ah[kk >> 1]);
            const uint4 *wrow = reinterpret_cast<const uint4 *>(wt + (k0 + kk) * NOUT);
            #pragma unroll
            for (int o8 = 0; o8 < NOUT / 8; ++o8) {
                const uint4 wv = wrow[o8];  // same address for the whole warp: broadcast
                const __half2 *wh = reinterpret_cast<const __half2 *>(&wv);
                #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 w = __half22float2(wh[j]);
                    acc[o8 * 8 + 2 * j] = fmaf(a, w.x, acc[o8 * 8 + 2 * j]);
                    acc[o8 * 8 + 2 * j + 1] = fmaf(a, w.y, acc[o8 * 8 + 2 * j + 1]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(TILE) mlp_fw_kernel(const __half *__restrict__ in, int in_stride,
                                                      const __half *__restrict__ weights, MlpShape s,
                                                      int64_t n, const int32_t *__restrict__ n_dev,
                                                      __half *__restrict__ hidden, __half *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half *wt = reinterpret_cast<__half *>(smem_raw);                 // transposed weights, all layers
    __half *act = wt + s.n_weights;                                    // [TILE][MAX_IN + ROWP]
    const int RS = MAX_IN + ROWP;
    const int tid = threadIdx.x;
    const int64_t n_alloc = n;
    n = b2n_eff_n(n, n_dev);
    // stage weights transposed: wt[l][k][o] = W_l[o][k]
    for (int l = 0; l < s.n_layers; ++l) {
        const int ni = layer_in(s, l), no = layer_out(s, l);
        for (int e = tid; e < ni * no; e += TILE) {
            const int o = e / ni, k = e - o * ni;
            wt[s.w_off[l] + k * no + o] = weights[s.w_off[l] + e];
        }
    }
    __syncthreads();
    __half *my = act + tid * RS;
    const int64_t n_tiles = (n + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * TILE;
        // coalesced tile load: in_width/8 16-byte chunks per row
        const int cpr = s.in_width / 8;
        __syncthreads();  // previous tile's rows are no longer read by their owners
        for (int e = tid; e < TILE * cpr; e += TILE) {
            const int r = e / cpr, c = e - r * cpr;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (row0 + r < n) v = *reinterpret_cast<const uint4 *>(in + (row0 + r) * in_stride + c * 8);
            *reinterpret_cast<uint4 *>(act + r * RS + c * 8) = v;
        }
        __syncthreads();
        const int64_t row = row0 + tid;
        for (int l = 0; l < s.n_hidden; ++l) {
            float acc[HID];
            #pragma unroll
            for (int o = 0; o < HID; ++o) acc[o] = 0.f;
            row_times_wt<HID>(my, layer_in(s, l), wt + s.w_off[l], acc);
            #pragma unroll
            for (int o = 0; o < HID; o += 8) {
                uint4 pk;
                __half2 *ph = reinterpret_cast<__half2 *>(&pk);
                #pragma unroll
                for (int j = 0; j < 4; ++j)
                    ph[j] = __floats2half2_rn(fmaxf(acc[o + 2 * j], 0.f), fmaxf(acc[o + 2 * j + 1], 0.f));
                *reinterpret_cast<uint4 *>(my + o) = pk;
                if (hidden != nullptr && row < n)
                    *reinterpret_cast<uint4 *>(hidden + ((int64_t)l * n_alloc + row) * HID + o) = pk;
            }
        }
        float acc[OUTW];
        #pragma unroll
        for (int o = 0; o < OUTW; ++o) acc[o] = 0.f;
        row_times_wt<OUTW>(my, HID, wt + s.w_off[s.n_hidden], acc);
        if (row < n) {
            #pragma unroll
            for (int o = 0; o < OUTW; o += 8) {
                uint4 pk;
                __half2 *ph = reinterpret_cast<__half2 *>(&pk);
                #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float a = acc[o + 2 * j], b = acc[o + 2 * j + 1];
                    if (s.out_act == 1) { a = 1.0f / (1.0f + __expf(-a)); b = 1.0f / (1.0f + __expf(-b)); }
                    ph[j] = __floats2half2_rn(a, b);
                }
                *reinterpret_cast<uint4 *>(out + row * OUTW + o) = pk;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward
// g_prev[k] = sum_o g[o] * W[o][k]  (W row-major (out,in) in smem; row o is a warp-wide broadcast)
template <int NOUT>
__device__ __forceinline__ void dgrad_row(const float *g, const __half *w, int n_in, float *gp) {
    for (int k = 0; k < n_in; ++k) gp[k] = 0.f;
    #pragma unroll
    for (int o = 0; o < NOUT; ++o) {
        const float go = g[o];
        const uint4 *wrow = reinterpret_cast<const uint4 *>(w + o * n_in);
        for (int k8 = 0; k8 < n_in / 8; ++k8) {
            const uint4 wv = wrow[k8];
            const __half2 *wh = reinterpret_cast<const __half2 *>(&wv);
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(wh[j]);
                gp[k8 * 8 + 2 * j] = fmaf(go, f.x, gp[k8 * 8 + 2 * j]);
                gp[k8 * 8 + 2 * j + 1] = fmaf(go, f.y, gp[k8 * 8 + 2 * j + 1]);
            }
        }
    }
}

// dW[o][k] += sum_rows G[row][o] * A[row][k] over the tile; G (TILE x NOUT) and A (TILE x n_in) in smem.
// Work split: thread -> (o, k-segment).  NOUT*n_in outputs over TILE threads.
template <int NOUT>
__device__ __forceinline__ void wgrad_tile(const __half *G, int gs, const __half *A, int as, int n_in,
                                           float *dW, int tid) {
    const int per = (NOUT * n_in) / TILE;       // outputs per thread: 32 (64x64), 40 (64x80), 16 (64x32), 8 (16x64)
    const int segs = n_in / per;                // k-segments per output row
    const int o = tid / segs, k0 = (tid - o * segs) * per;
    float acc[40];
    #pragma unroll
    for (int j = 0; j < 40; ++j) acc[j] = 0.f;
    for (int r = 0; r < TILE; ++r) {
        const float g = __half2float(G[r * gs + o]);
        const __half *a = A + r * as + k0;
        #pragma unroll
        for (int j8 = 0; j8 < 5; ++j8) {
            if (j8 * 8 < per) {
                const uint4 av = *reinterpret_cast<const uint4 *>(a + j8 * 8);
                const __half2 *ah = reinterpret_cast<const __half2 *>(&av);
                #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(ah[j]);
                    acc[j8 * 8 + 2 * j] = fmaf(g, f.x, acc[j8 * 8 + 2 * j]);
                    acc[j8 * 8 + 2 * j + 1] = fmaf(g, f.y, acc[j8 * 8 + 2 * j + 1]);
                }
            }
        }
    }
    #pragma unroll
    for (int j = 0; j < 40; ++j)
        if (j < per) dW[o * n_in + k0 + j] += acc[j];
}

__global__ void __launch_bounds__(TILE) mlp_bw_kernel(const __half *__restrict__ dy, const __half *__restrict__ in,
                                                      int in_stride, const __half *__restrict__ weights,
                                                      MlpShape s, int64_t n, const int32_t *__restrict__ n_dev,
                                                      const __half *__restrict__ hidden,
                                                      const __half *__restrict__ out, float grad_scale,
                                                      __half *__restrict__ din, float *__restrict__ grad_weights) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int RS = MAX_IN + ROWP, GS = HID + ROWP;
    float *dW = reinterpret_cast<float *>(smem_raw);                     // fp32 partial sums, all layers
    __half *w = reinterpret_cast<__half *>(dW + s.n_weights);            // weights row-major (out,in)
    __half *A = w + s.n_weights;                                         // activation tile [TILE][RS]
    __half *G = A + TILE * RS;                                           // gradient tile   [TILE][GS]
    const int tid = threadIdx.x;
    const int64_t n_alloc = n;
    n = b2n_eff_n(n, n_dev);
    for (int e = tid; e < s.n_weights; e += TILE) { w[e] = weights[e]; dW[e] = 0.f; }
    __syncthreads();
    const int64_t n_tiles = (n + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * TILE, row = row0 + tid;
        const bool live = row < n;
        // gradient at the (activated) output, this thread's row
        float g[HID];
        {
            uint4 dv[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)}, ov[2] = {dv[0], dv[0]};
            if (live) {
                dv[0] = *reinterpret_cast<const uint4 *>(dy + row * OUTW);
                dv[1] = *reinterpret_cast<const uint4 *>(dy + row * OUTW + 8);
                if (s.out_act == 1) {
                    ov[0] = *reinterpret_cast<const uint4 *>(out + row * OUTW);
                    ov[1] = *reinterpret_cast<const uint4 *>(out + row * OUTW + 8);
                }
            }
            const __half *dh = reinterpret_cast<const __half *>(dv), *oh = reinterpret_cast<const __half *>(ov);
            #pragma unroll
            for (int o = 0; o < OUTW; ++o) {
                float v = __half2float(dh[o]);
                if (s.out_act == 1) { const float y = __half2float(oh[o]); v *= y * (1.0f - y); }
                g[o] = v;
            }
        }
        int g_width = OUTW;
        for (int l = s.n_hidden; l >= 0; --l) {
            const int ni = layer_in(s, l);
            // stage this layer's input activations (tile) and the gradient rows
            __syncthreads();
            {
                const __half *src = (l == 0) ? in : hidden + (int64_t)(l - 1) * n_alloc * HID;
                const int stride = (l == 0) ? in_stride : HID;
                const int cpr = ni / 8;
                for (int e = tid; e < TILE * cpr; e += TILE) {
                    const int r = e / cpr, c = e - r * cpr;
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (row0 + r < n) v = *reinterpret_cast<const uint4 *>(src + (row0 + r) * stride + c * 8);
                    *reinterpret_cast<uint4 *>(A + r * RS + c * 8) = v;
                }
                #pragma unroll
                for (int o = 0; o < HID; o += 2)
                    if (o < g_width)
                        *reinterpret_cast<__half2 *>(G + tid * GS + o) = __floats2half2_rn(g[o], g[o + 1]);
            }
            __syncthreads();
            // weight gradient of layer l
            if (g_width == OUTW) wgrad_tile<OUTW>(G, GS, A, RS, ni, dW + s.w_off[l], tid);
            else                 wgrad_tile<HID>(G, GS, A, RS, ni, dW + s.w_off[l], tid);
            // gradient w.r.t. this layer's input (skipped for layer 0 unless requested)
            if (l > 0 || din != nullptr) {
                float gp[MAX_IN];
                if (g_width == OUTW) dgrad_row<OUTW>(g, w + s.w_off[l], ni, gp);
                else                 dgrad_row<HID>(g, w + s.w_off[l], ni, gp);
                if (l > 0) {
                    // ReLU': the saved post-activation is > 0 exactly where the pre-activation was
                    const __half *a = A + tid * RS;
                    #pragma unroll
                    for (int k = 0; k < HID; ++k) g[k] = (__half2float(a[k]) > 0.f) ? gp[k] : 0.f;
                    g_width = HID;
                } else if (live) {
                    for (int k = 0; k < ni; k += 2)
                        *reinterpret_cast<__half2 *>(din + row * in_stride + k) = __floats2half2_rn(gp[k], gp[k + 1]);
                }
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < s.n_weights; e += TILE) {
        const float v = dW[e] * grad_scale;
        if (v != 0.f) atomicAdd(grad_weights + e, v);
    }
}

extern "C" int b2n_mlp_fw(const b2n_half *in, int in_stride, int in_width, const b2n_half *weights, int n_hidden,
                          int output_activation, int64_t n, const int32_t *n_dev, b2n_half *hidden,
                          b2n_half *out, void *stream) {
    MlpShape s;
    if (make_shape(in_width, n_hidden, output_activation, s)) return 1;
    B2N_CHECK_ARG(in_stride >= in_width && in_stride % 8 == 0, "in_stride must be >= in_width and a multiple of 8");
    if (n <= 0) return 0;
    const size_t smem = (size_t)s.n_weights * 2 + (size_t)TILE * (MAX_IN + ROWP) * 2;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(mlp_fw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        attr_set = true;
    }
    mlp_fw_kernel<<<b2n_grid((n + TILE - 1) / TILE, 4), TILE, smem, (cudaStream_t)stream>>>(
        (const __half *)in, in_stride, (const __half *)weights, s, n, n_dev, (__half *)hidden, (__half *)out);
    B2N_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2n_mlp_bw(const b2n_half *dL_dout, const b2n_half *in, int in_stride, int in_width,
                          const b2n_half *weights, int n_hidden, int output_activation, int64_t n,
                          const int32_t *n_dev, const b2n_half *hidden, const b2n_half *out, float grad_scale,
                          b2n_half *dL_din, float *grad_weights, void *stream) {
    MlpShape s;
    if (make_shape(in_width, n_hidden, output_activation, s)) return 1;
    B2N_CHECK_ARG(in_stride >= in_width && in_stride % 8 == 0, "in_stride must be >= in_width and a multiple of 8");
    B2N_CHECK_ARG(hidden != nullptr && grad_weights != nullptr, "hidden activations and grad_weights are required");
    B2N_CHECK_ARG(output_activation == 0 || out != nullptr, "sigmoid backward needs the forward output");
    if (n <= 0) return 0;
    const size_t smem = (size_t)s.n_weights * 6 + (size_t)TILE * (MAX_IN + ROWP) * 2 + (size_t)TILE * (HID + ROWP) * 2;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(mlp_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr_set = true;
    }
    mlp_bw_kernel<<<b2n_grid((n + TILE - 1) / TILE, 2), TILE, smem, (cudaStream_t)stream>>>(
        (const __half *)dL_dout, (const __half *)in, in_stride, (const __half *)weights, s, n, n_dev,
        (const __half *)hidden, (const __half *)out, grad_scale, (__half *)dL_din, grad_weights);
    B2N_LAUNCH_CHECK();
    return 0;
}
