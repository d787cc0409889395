// tc_common.cuh -- thin inline-PTX layer over the Blackwell tensor-core path used by field_tc.cu:
// tcgen05.mma (kind::f16, cta_group::1, operands from shared memory, fp32 accumulators in TMEM),
// tcgen05.alloc/ld/commit, mbarriers, the async-proxy fence and cp.async.bulk (TMA bulk copy).
// sm_100a only; nothing here is portable and nothing is meant to be.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory matrix descriptor, no swizzle ("interleave") canonical layout ---------------
// Core matrix = 8 rows x 16 bytes, stored as 128 contiguous bytes.  For a K-major operand (rows = M/N index,
// 16 B = 8 consecutive K elements): SBO = byte stride between 8-row groups, LBO = byte stride between
// 16-byte K chunks.  The same bytes read as an MN-major operand (rows = K index) swap the two roles.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version 1 (sm_100)
    // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
    return d;
}

// instruction descriptor for kind::f16: fp16 x fp16 -> fp32, dense
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4)                        // c_format = F32
           | (0u << 7) | (0u << 10)         // a_format = b_format = F16
           | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16)
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// arrive on an mbarrier when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM -------------------------------------------------------------------------------------
// executed by one full warp; writes the TMEM base address (lane 0, column c) into *dst_smem
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// 32 lanes x 32 bit, 16 consecutive columns: thread t of the warp receives lane (base_lane + t), columns c..c+15
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    #pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 consecutive columns with ONE tcgen05.ld and one wait
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    #pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// zero 16 consecutive columns of this warp's 32 lanes (tcgen05.st), completion via tmem_st_wait()
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
    const uint32_t z = 0;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z)
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// four 16-column loads in flight, one wait: v[0..63] = columns c..c+63 of this thread's lane
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float *v) {
    uint32_t r[64];
    #pragma unroll
    for (int q = 0; q < 4; ++q) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[16 * q + 0]), "=r"(r[16 * q + 1]), "=r"(r[16 * q + 2]), "=r"(r[16 * q + 3]), "=r"(r[16 * q + 4]),
              "=r"(r[16 * q + 5]), "=r"(r[16 * q + 6]), "=r"(r[16 * q + 7]), "=r"(r[16 * q + 8]), "=r"(r[16 * q + 9]),
              "=r"(r[16 * q + 10]), "=r"(r[16 * q + 11]), "=r"(r[16 * q + 12]), "=r"(r[16 * q + 13]),
              "=r"(r[16 * q + 14]), "=r"(r[16 * q + 15])
            : "r"(taddr + 16 * q)
            : "memory");
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    #pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- coalesced copies between a canonical tile and row-major global memory ---------------------------
// NCH = 16-byte chunks per row.  Thread mapping: consecutive threads take consecutive chunks of a row, so every
// warp-level global access covers whole 128-byte lines, and (thanks to the ACT_LBO padding) the shared-memory
// side is conflict free.  rows_valid = number of rows of this tile that exist in global memory.
template <int NCH>
__device__ __forceinline__ void tile_load(unsigned char *tile, const __half *g_row0, int64_t rows_valid, int tid) {
    #pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int j = i * 128 + tid, row = j / NCH, c = j % NCH;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (row < rows_valid) v = __ldg(reinterpret_cast<const uint4 *>(g_row0 + (int64_t)row * (NCH * 8)) + c);
        *reinterpret_cast<uint4 *>(tile + (uint32_t)c * 2064u + (uint32_t)row * 16) = v;
    }
}
template <int NCH>
__device__ __forceinline__ void tile_store(const unsigned char *tile, __half *g_row0, int64_t rows_valid, int tid) {
    #pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int j = i * 128 + tid, row = j / NCH, c = j % NCH;
        const uint4 v = *reinterpret_cast<const uint4 *>(tile + (uint32_t)c * 2064u + (uint32_t)row * 16);
        if (row < rows_valid) *(reinterpret_cast<uint4 *>(g_row0 + (int64_t)row * (NCH * 8)) + c) = v;
    }
}

// ---- cp.async (LDGSTS): asynchronous 16-byte global -> shared copies, zero-filled when !valid -----------
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem, bool valid) {
    const uint32_t sz = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int NCH>
__device__ __forceinline__ void tile_load_async(unsigned char *tile, const __half *g_row0, int64_t rows_valid, int tid) {
    #pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int j = i * 128 + tid, row = j / NCH, c = j % NCH;
        const bool ok = row < rows_valid;
        const __half *src = g_row0 + (ok ? ((int64_t)row * (NCH * 8) + c * 8) : 0);
        cp_async16(tile + (uint32_t)c * 2064u + (uint32_t)row * 16, src, ok);
    }
}

// gather variant: row r of the tile comes from global row row_idx[r] (row_idx: 128 ints in shared memory)
template <int NCH>
__device__ __forceinline__ void tile_gather_async(unsigned char *tile, const __half *g_base, const int32_t *row_idx,
                                                  int64_t rows_valid, int tid) {
    #pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int j = i * 128 + tid, row = j / NCH, c = j % NCH;
        const bool ok = row < rows_valid;
        const __half *src = g_base + (ok ? ((int64_t)row_idx[row] * (NCH * 8) + c * 8) : 0);
        cp_async16(tile + (uint32_t)c * 2064u + (uint32_t)row * 16, src, ok);
    }
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// ---- TMA bulk copy global -> shared (1-D, 16-byte granularity), completion on an mbarrier ---------
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- canonical (no-swizzle) tile addressing ------------------------------------------------------
// activation tile [128 rows x K]: 16-byte chunk c of row r lives at  c*ACT_LBO + (r/8)*128 + (r%8)*16 = c*ACT_LBO + r*16.
// ACT_LBO carries 16 bytes of padding per chunk plane so that BOTH access patterns are bank-conflict free:
// "thread = row, fixed chunk" (epilogue) and "8 threads = the 8 chunks of one row" (coalesced global copies).
constexpr uint32_t ACT_SBO = 128;    // between 8-row groups
constexpr uint32_t ACT_LBO = 2064;   // between 16-byte K chunks (16 row groups x 128 B + 16 B pad)
constexpr uint32_t TILE16_BYTES = 2 * ACT_LBO, TILE32_BYTES = 4 * ACT_LBO, TILE64_BYTES = 8 * ACT_LBO;
__device__ __forceinline__ uint32_t act_off(int r, int chunk) {
    return (uint32_t)chunk * ACT_LBO + (uint32_t)(r >> 3) * ACT_SBO + (uint32_t)(r & 7) * 16;
}
// weight matrix [N rows x K] (row-major (out,in) = K-major B operand): chunk c of row n at
// c*(N/8)*128 + (n/8)*128 + (n%8)*16;  half index = that / 2 + (k % 8)
__host__ __device__ constexpr uint32_t w_off_halves(int n, int k, int N) {
    return (uint32_t)((k >> 3) * (N >> 3) + (n >> 3)) * 64 + (uint32_t)(n & 7) * 8 + (uint32_t)(k & 7);
}

}  // namespace tc
