// peer.cu -- data-parallel exchange over NVLink peer memory (one process per GPU, one node).
// Replaces what the reference gets from DistributedDataParallel + NCCL (ngp_pl/train.py:197-208: DDPPlugin,
// one gradient all-reduce per step, every rank runs the whole optimiser): here the gradient reduction, the
// optimiser and the parameter broadcast are ONE kernel.
//
// Design (DESIGN.md "Multi-GPU"): every rank keeps its gradient vector g (fp32) and its fp16 working copy h of the
// parameters in a cudaMalloc block whose CUDA-IPC handle the other ranks have opened, so all ranks can address all
// g_r and h_r.  The flat parameter vector is sharded: rank s owns elements [s*shard, (s+1)*shard) of the fp32 master
// copy and of Adam's moments.  After the backward pass:
//   barrier  -- every rank's g is complete
//   adam_peer: the owner READS its slice of every rank's g over NVLink (P2P loads, summed in rank order), runs Adam on
//              its master slice and WRITES the fp16 result into every rank's h (P2P stores) -- reduce-scatter,
//              optimiser and all-gather without an intermediate buffer or a second pass over HBM
//   barrier  -- every rank's h is complete and nobody still reads g
// after which each rank zeroes its own g.  The barrier is a flag exchange through the same peer mappings: every rank
// stores a monotonically increasing epoch into its slot on every peer (st.release.sys after a system fence) and
// spins on its local slots (ld.acquire.sys) with a wall-clock bound, so a dead peer raises an error flag instead of
// hanging the GPU.  All of it is plain kernels on the caller's stream: CUDA-graph capturable, no communicator.
#include "common.cuh"

#define PEER_MAX_WORLD 16

// ------------------------------------------------------------------------------------------------ memory
extern "C" int b2n_peer_alloc(int64_t bytes, void **ptr, void *handle64) {
    B2N_CHECK_ARG(bytes > 0 && ptr != nullptr && handle64 != nullptr, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
    if (e != cudaSuccess) {
        b2n_set_error("b2n_peer_alloc: cudaMalloc: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return 2;
    }
    e = cudaMemset(*ptr, 0, (size_t)bytes);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, *ptr);
    if (e != cudaSuccess) {
        b2n_set_error("b2n_peer_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
        cudaFree(*ptr); *ptr = nullptr;
        (void)cudaGetLastError();
        return 2;
    }
    return 0;
}

extern "C" int b2n_peer_open(const void *handle64, void **ptr) {
    B2N_CHECK_ARG(handle64 != nullptr && ptr != nullptr, "bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    const cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        b2n_set_error("b2n_peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();                           // reported through the status; do not leave it pending
        return 2;
    }
    return 0;
}

extern "C" int b2n_peer_close(void *ptr) {
    if (ptr == nullptr) return 0;
    const cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) { b2n_set_error("b2n_peer_close: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return 2; }
    return 0;
}

extern "C" int b2n_peer_free(void *ptr) {
    if (ptr == nullptr) return 0;
    const cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) { b2n_set_error("b2n_peer_free: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return 2; }
    return 0;
}

// ------------------------------------------------------------------------------------------------ barrier
struct PeerPtrs {
    void *p[PEER_MAX_WORLD];
};

__device__ __forceinline__ void st_release_sys(uint32_t *addr, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *addr) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// flags: PEER_MAX_WORLD words per rank, word r = last epoch rank r has announced to this rank.
// state (local): [0] = epoch of the last completed barrier, [1] = error flag (sticky).
__global__ void __launch_bounds__(32) peer_barrier_kernel(const __grid_constant__ PeerPtrs flags, int rank, int world,
                                                          uint32_t *state, uint64_t timeout_ns) {
    const int r = threadIdx.x;
    const uint32_t epoch = state[0] + 1;
    __syncwarp();
    if (r < world) {
        __threadfence_system();                              // everything this GPU wrote so far (also to peers) first
        st_release_sys(reinterpret_cast<uint32_t *>(flags.p[r]) + rank, epoch);
        const uint32_t *mine = reinterpret_cast<const uint32_t *>(flags.p[rank]) + r;
        const uint64_t t0 = global_ns();
        // epochs only grow; a peer that is already one barrier ahead has announced a larger value
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
            if (global_ns() - t0 > timeout_ns) { state[1] = 1u + (uint32_t)r; break; }
            __nanosleep(64);
        }
    }
    __syncwarp();
    if (r == 0) state[0] = epoch;
}

extern "C" int b2n_peer_barrier(void *const *flag_ptrs, int rank, int world, uint32_t *state, double timeout_s,
                                void *stream) {
    B2N_CHECK_ARG(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, "bad rank / world");
    PeerPtrs f;
    for (int r = 0; r < PEER_MAX_WORLD; ++r) f.p[r] = r < world ? flag_ptrs[r] : nullptr;
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, rank, world, state, (uint64_t)(timeout_s * 1e9));
    B2N_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------ fused optimiser
// Optional fp16 wire format for the hash-table gradients: each rank converts the [lo, hi) part of its (loss-scaled) fp32
// gradient vector to fp16 once (saturating), clearing that part of the fp32 vector in the same pass, and the owners
// read the fp16 copies -- half the NVLink bytes of the dominant term.  tiny-cuda-nn itself keeps per-rank table
// gradients in fp16 at the same loss scale, so this loses nothing against the reference's DDP.  The few MLP weights
// outside [lo, hi) stay fp32 on the wire (Adam turns their rounding into visible trajectory differences).
__global__ void __launch_bounds__(256) grad_pack_half_kernel(float4 *__restrict__ g, uint2 *__restrict__ g16, int64_t lo4,
                                                             int64_t hi4) {
    for (int64_t i = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = g[i];
        v.x = fminf(fmaxf(v.x, -65504.f), 65504.f); v.y = fminf(fmaxf(v.y, -65504.f), 65504.f);
        v.z = fminf(fmaxf(v.z, -65504.f), 65504.f); v.w = fminf(fmaxf(v.w, -65504.f), 65504.f);
        const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t *>(&a);
        o.y = *reinterpret_cast<const uint32_t *>(&b);
        g16[i] = o;
        g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

extern "C" int b2n_grad_pack_half(float *grad, b2n_half *grad16, int64_t lo, int64_t hi, void *stream) {
    B2N_CHECK_ARG(lo % 4 == 0 && hi % 4 == 0 && lo <= hi, "range bounds must be multiples of 4");
    if (lo == hi) return 0;
    grad_pack_half_kernel<<<b2n_grid(b2n_blocks((hi - lo) / 4, 256), 8), 256, 0, (cudaStream_t)stream>>>(
        (float4 *)grad, (uint2 *)grad16, lo / 4, hi / 4);
    B2N_LAUNCH_CHECK();
    return 0;
}

// Per owned parameter: W gradient reads (W-1 of them over NVLink), p/m/v read + written locally, W fp16 writes
// (W-1 over NVLink).  fp32 gradients are left as they are: each rank clears its own vector with one memset after the
// closing barrier (a local HBM fill is cheaper than W-1 remote zero stores per element).
template <bool HALF_GRAD>
__global__ void __launch_bounds__(256) adam_peer_kernel(float4 *__restrict__ p, float4 *__restrict__ m,
                                                        float4 *__restrict__ v, const __grid_constant__ PeerPtrs g,
                                                        const __grid_constant__ PeerPtrs g16, int64_t half_lo4,
                                                        int64_t half_hi4,
                                                        const __grid_constant__ PeerPtrs h, int world, int64_t first4,
                                                        int64_t n4, float lr, float b1, float b2, float eps,
                                                        float inv_scale, int step, const b2n_hyper *__restrict__ hyper,
                                                        const __grid_constant__ PeerPtrs hy,
                                                        const uint32_t *__restrict__ barrier_state) {
    // a timed-out barrier (sticky error word) means some peer's gradients are incomplete: leave everything untouched
    if (barrier_state != nullptr && barrier_state[1] != 0u) return;
    if (hyper != nullptr) {
        // GradScaler semantics: any rank's overflow skips the step.  ONE thread per CTA looks at the peers' flags (a
        // volatile load per thread would put more sectors on NVLink than the gradients themselves).
        __shared__ int s_bad;
        if (threadIdx.x == 0) {
            int bad = hyper->found_inf;
            for (int r = 0; r < world; ++r)
                if (hy.p[r] != nullptr) bad |= reinterpret_cast<const volatile b2n_hyper *>(hy.p[r])->found_inf;
            s_bad = bad;
        }
        __syncthreads();
        if (s_bad) return;                                   // (the gradient vectors are cleared by the caller's memset)
        lr = hyper->lr;
        step = hyper->step - hyper->skipped;
        if (hyper->loss_scale > 0.f) inv_scale /= hyper->loss_scale;
    }
    const float c1 = 1.0f - powf(b1, (float)step), c2 = 1.0f - powf(b2, (float)step);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool use16 = HALF_GRAD && first4 + i >= half_lo4 && first4 + i < half_hi4;
        #pragma unroll 8
        for (int r = 0; r < world; ++r) {                    // fixed order: the sum does not depend on timing
            if (use16) {
                const uint2 t = reinterpret_cast<const uint2 *>(g16.p[r])[first4 + i];
                const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&t.x));
                const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&t.y));
                gg.x += a.x; gg.y += a.y; gg.z += b.x; gg.w += b.y;
            } else {
                const float4 t = reinterpret_cast<const float4 *>(g.p[r])[first4 + i];
                gg.x += t.x; gg.y += t.y; gg.z += t.z; gg.w += t.w;
            }
        }
        float4 pp = p[i], mm = m[i], vv = v[i];
        float *P = &pp.x, *G = &gg.x, *M = &mm.x, *V = &vv.x;
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gr = G[k] * inv_scale;
            M[k] = b1 * M[k] + (1.0f - b1) * gr;
            V[k] = b2 * V[k] + (1.0f - b2) * gr * gr;
            P[k] -= lr * (M[k] / c1) / (sqrtf(V[k] / c2) + eps);
        }
        p[i] = pp; m[i] = mm; v[i] = vv;
        const __half2 h0 = __floats2half2_rn(pp.x, pp.y), h1 = __floats2half2_rn(pp.z, pp.w);
        uint2 packed;
        packed.x = *reinterpret_cast<const uint32_t *>(&h0);
        packed.y = *reinterpret_cast<const uint32_t *>(&h1);
        for (int r = 0; r < world; ++r) reinterpret_cast<uint2 *>(h.p[r])[first4 + i] = packed;
    }
}

extern "C" int b2n_adam_step_peer(float *param_shard, float *exp_avg, float *exp_avg_sq, void *const *grad_ptrs,
                                  void *const *grad16_ptrs, int64_t half_lo, int64_t half_hi, void *const *half_ptrs,
                                  int world, int64_t shard_first, int64_t shard_n, float lr, float beta1, float beta2,
                                  float eps, float inv_scale, int step, const b2n_hyper *hyper_dev,
                                  void *const *hyper_ptrs, const uint32_t *barrier_state, void *stream) {
    B2N_CHECK_ARG(world >= 1 && world <= PEER_MAX_WORLD, "bad world size");
    B2N_CHECK_ARG(shard_n % 4 == 0 && shard_first % 4 == 0, "shard bounds must be multiples of 4");
    B2N_CHECK_ARG(half_lo % 4 == 0 && half_hi % 4 == 0, "fp16 range bounds must be multiples of 4");
    B2N_CHECK_ARG(step >= 1 || hyper_dev != nullptr, "step is 1-based");
    if (shard_n == 0) return 0;
    PeerPtrs g, g16, h, hy;
    for (int r = 0; r < PEER_MAX_WORLD; ++r) {
        hy.p[r] = (r < world && hyper_ptrs != nullptr) ? hyper_ptrs[r] : nullptr;
        g.p[r] = r < world ? grad_ptrs[r] : nullptr;
        g16.p[r] = (r < world && grad16_ptrs != nullptr) ? grad16_ptrs[r] : nullptr;
        h.p[r] = r < world ? half_ptrs[r] : nullptr;
    }
    const unsigned grid = b2n_grid(b2n_blocks(shard_n / 4, 256), 32);     // like adam_kernel: ~4 waves beat a persistent grid
    if (grad16_ptrs != nullptr && half_hi > half_lo)
        adam_peer_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(
            (float4 *)param_shard, (float4 *)exp_avg, (float4 *)exp_avg_sq, g, g16, half_lo / 4, half_hi / 4, h, world,
            shard_first / 4, shard_n / 4, lr, beta1, beta2, eps, inv_scale, step, hyper_dev, hy, barrier_state);
    else
        adam_peer_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(
            (float4 *)param_shard, (float4 *)exp_avg, (float4 *)exp_avg_sq, g, g16, 0, 0, h, world, shard_first / 4,
            shard_n / 4, lr, beta1, beta2, eps, inv_scale, step, hyper_dev, hy, barrier_state);
    B2N_LAUNCH_CHECK();
    return 0;
}
