// sm_100a primitives for the tensor-core path: mbarrier, tcgen05 (TMEM alloc, MMA
// kind::tf32, commit, ld), bulk async copy (TMA engine, 1-D), shared-memory matrix and
// instruction descriptors.  Inline PTX only; no CUTLASS/CuTe dependency.
//
// Operand layout used everywhere here: K-major, SWIZZLE_128B.  One "K block" holds
// 32 tf32 (128 bytes) of K for `rows` rows: row r starts at (r/8)*1024 + (r%8)*128 and
// its eight 16-byte chunks are stored at chunk index (c ^ (r%8)).  A tcgen05.mma of
// kind::tf32 consumes K=8 (32 bytes); stepping K inside a block advances the
// descriptor start address by 32 bytes, stepping to the next block by rows*128 bytes.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mft {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t count) {   // count arrivals at once
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes or
// the hint (ns) expires -- a parked thread takes no issue slots, unlike a polling loop.  This
// matters here: the single-thread MMA and TMA warps have the highest warp ids of the CTA and the
// scheduler favours high ids, so a spinning waiter starves the epilogue warp on its sub-partition.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(200000u)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {   // ~2 s at 2 GHz
            printf("mft: mbarrier wait timed out (block %d thread %d parity %u)\n", (int)blockIdx.x,
                   (int)threadIdx.x, parity);
            __trap();
        }
    }
}

// ------------------------------------------------------------------ fences
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ------------------------------------------------------------------ TMEM
// All 32 lanes of ONE warp call these (sync.aligned).  ncols: power of two >= 32.
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 32 consecutive columns -> 32 registers per thread (thread i <-> lane base+i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 16 consecutive columns -> 16 registers per thread.
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 32 consecutive columns <- 32 registers per thread (thread i <-> lane base+i): the A operand of a
// TMEM-sourced MMA (row = lane, column = k as a 32-bit word) is written this way.
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ MMA
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by ONE thread.
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem: 128 lanes x 8 columns of 32-bit words] * B[smem desc], kind::tf32, issued by ONE thread.
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    const uint32_t z = 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
// mbarrier arrives once every MMA issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}

// Instruction descriptor, kind::tf32: D = F32, A = B = TF32, both K-major (or MN-major when flagged).
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool a_mn_major = false,
                                                       bool b_mn_major = false) {
    return (1u << 4)                        // c_format  = F32
           | (2u << 7)                      // a_format  = TF32
           | (2u << 10)                     // b_format  = TF32
           | ((a_mn_major ? 1u : 0u) << 15) // a_major
           | ((b_mn_major ? 1u : 0u) << 16) // b_major
           | ((uint32_t)(N >> 3) << 17)     // n_dim
           | ((uint32_t)(M >> 4) << 24);    // m_dim
}

// Shared-memory matrix descriptor, SWIZZLE_128B, 8-row groups `sbo_bytes` apart
// (1024 for the dense K-major layout above); `lbo_bytes` only matters for MN-major.
//
// layout_type 2 = SWIZZLE_128B (16-byte chunks XOR row%8; K-major operands here);
// layout_type 1 = SWIZZLE_128B_BASE32B (32-byte chunks XOR row%4): the only swizzled layout the
// hardware accepts for MN-major 32-bit (tf32) operands -- atoms are 4 k-rows x 128 bytes.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t lbo_bytes,
                                                    uint32_t layout_type = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;     // leading byte offset, bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;     // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;                      // layout type
    return d;
}

// ------------------------------------------------------------------ bulk copy (TMA engine, 1-D)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ------------------------------------------------------------------ TMA (tensor map) 2-D tile load
// Box {inner = c0.., outer = c1..} of the mapped row-major matrix -> shared memory, swizzled as the
// tensor map says; completes `bytes of the box` on the mbarrier.  tmap: address of a CUtensorMap in
// kernel-parameter (__grid_constant__) or global memory.
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(smem_u32(dst_smem)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// ------------------------------------------------------------------ CTA pairs (cta_group::2)
// Two CTAs of a cluster on one TPC issue ONE tcgen05.mma of M = 256: each CTA supplies its own 128 rows of
// A and HALF of the N rows of B from the same shared-memory offsets, and receives its 128 x N slice of D in
// its own TMEM.  Only the leader (cluster rank 0) issues; tcgen05.commit multicasts the completion to the
// barrier at the same offset in both CTAs; the peer's warps arrive on the leader's barriers through the
// cluster shared-memory window (mapa).
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {     // every thread of both CTAs
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `bar` (a shared-memory object of THIS CTA) as seen in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(200000u)
        : "memory");
    return ok != 0;
}
// waits on a barrier whose arrivals may come from the peer CTA (acquire at cluster scope); bounded like mbar_wait
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("mft: cluster mbarrier wait timed out (block %d thread %d parity %u)\n", (int)blockIdx.x,
                   (int)threadIdx.x, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {   // one warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_tf32_ss2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    const uint32_t z = 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs once every MMA issued so far has completed
__device__ __forceinline__ void mma_commit2(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(mask)
        : "memory");
}

// Register re-balancing between warp roles (all warps of a warpgroup execute it).
template <int N> __device__ __forceinline__ void reg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N> __device__ __forceinline__ void reg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

// ------------------------------------------------------------------ misc
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// Operand rounding in the producers: tcgen05 kind::tf32 reads only the upper 19 bits of an fp32
// operand, so round-to-nearest (ties away, what cvt.rna does) is ONE integer add of half an ulp --
// the three-instruction cvt.rna expansion (inf test, add, mask) made up a fifth of the producer loop.
// Differs from cvt.rna only for |x| >= 2^128 * (1 - 2^-12) and non-finite values, which the
// normalised activations / gradients of this path never are (a NaN stays a NaN).
__device__ __forceinline__ float to_tf32_fast(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// float offset of element (row, k_in_block) inside one K block of the layout above
__device__ __forceinline__ int sw128_offset(int row, int k) {
    return (row >> 3) * 256 + (row & 7) * 32 + ((((k >> 2) ^ (row & 7)) << 2) | (k & 3));
}

}  // namespace umma
}  // namespace mft
