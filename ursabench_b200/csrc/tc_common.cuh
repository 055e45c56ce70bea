// tcgen05 / TMEM / TMA building blocks shared by the tensor-core BMA kernels (sm_100a).
//   SASS: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG, tcgen05.commit -> UTCBAR
#pragma once
#include <cuda.h>

#include "async.cuh"
#include "common.cuh"

namespace ursa {

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the mbarrier receives one arrival once all tcgen05.mma previously issued by this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar_addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// the same load without the wait: issue several, then ONE tmem_wait_ld() -- the loads' latencies overlap
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// ReLU that PROPAGATES NaN: fmaxf(NaN, 0) = 0 would turn an FP16 range overflow (inf - inf = NaN in the accumulators of the
// FP16-split stage kernels) into silently wrong, finite activations.  max.NaN returns NaN if either operand is NaN.
__device__ __forceinline__ float relu_nan(float x) {
    float r;
    asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rn_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// raw-address mbarrier / TMA helpers (tile rings are addressed by 32-bit shared-window offsets)
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug becomes a trap (launch error) instead of a hung GPU
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait_a(bar, parity); ++spin)
        if (spin > (1u << 22)) __trap();   // ~seconds: no legitimate wait in these kernels exceeds milliseconds
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d_a(uint32_t dst, const CUtensorMap *tmap, int x, int y, int z, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_a(uint32_t dst, const CUtensorMap *tmap, int x, int y, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_a(uint32_t dst, const CUtensorMap *tmap, int c0, int c1, int c2, int c3,
                                              uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d_a(uint32_t dst, const CUtensorMap *tmap, int c0, int c1, int c2, int c3, int c4,
                                              uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar)
        : "memory");
}

// ---- CTA pair (cta_group::2): two CTAs of a cluster on the two SMs of a TPC run ONE M = 256 MMA; each streams its own A half
// and HALF of B from its shared memory.  Only the leader (cluster rank 0) issues the MMAs; both CTAs load by TMA and signal the
// leader's barrier; the leader's commits are multicast to the same barrier offset in both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// one arrival on the barrier at this shared-memory offset in every CTA of `cta_mask` once the prior MMAs have completed
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar_addr, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     bar_addr),
                 "h"(cta_mask)
                 : "memory");
}
// TMA loads whose completion is counted on a barrier of the PEER CTA (cluster addresses for destination and barrier)
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap *tmap, int x, int y, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_2cta(uint32_t dst, const CUtensorMap *tmap, int c0, int c1, int c2, int c3,
                                                 uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor) for rows of ROW_BYTES (= swizzle span):
// start>>4 | LBO = 1 (unused for swizzled K-major) | SBO = 8 rows * ROW_BYTES | version 1 (sm_100) |
// layout type 2 (SWIZZLE_128B) or 4 (SWIZZLE_64B)
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
    static_assert(ROW_BYTES == 128 || ROW_BYTES == 64, "row = one swizzle span");
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * ROW_BYTES) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(ROW_BYTES == 128 ? 2 : 4) << 61;
    return d;
}

// cute::UMMA::InstrDescriptor for kind::tf32: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), K-major both,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28
__device__ __forceinline__ uint32_t make_tf32_idesc(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- host: tensor-map encoding through the driver entry point (no libcuda link) --------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// tensor of `rank` dims (dims[0] innermost, strides in BYTES for dims 1..rank-1), box in elements,
// swizzle_bytes in {0 (none), 64, 128}; out-of-bounds elements read as zero.  f16 != 0: elements are halves, else fp32.
inline int make_tensor_map_t(CUtensorMap *tm, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                             const uint32_t *box, int swizzle_bytes, int f16) {
    EncodeTiledFn fn = tensor_map_encoder();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return URSA_ERR_CUDA;
    }
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; ++i) {
        d[i] = dims[i];
        b[i] = box[i];
        e[i] = 1;
        if (i + 1 < rank) s[i] = strides_bytes[i];
    }
    CUresult r = fn(tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                    const_cast<void *>(base), d, s, b, e,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                         : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu ..] box [%u %u %u ..]", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                  box[0], box[1], rank > 2 ? box[2] : 0);
        return URSA_ERR_CUDA;
    }
    return URSA_OK;
}
inline int make_tensor_map(CUtensorMap *tm, const float *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                           const uint32_t *box, int swizzle_bytes) {
    return make_tensor_map_t(tm, base, rank, dims, strides_bytes, box, swizzle_bytes, 0);
}

// ---- TMA tensor STORE (shared -> global) in the bulk async-group ------------------------------------------------------
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tmap, uint32_t src_smem, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(src_smem),
                 "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING their shared-memory source (it may be overwritten)
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace ursa
