// PTX helpers shared by the tcgen05 kernels: mbarrier, cp.async / bulk copy, UMMA descriptors,
// tcgen05.mma / commit / ld.  Bit layouts follow cute/arch/mma_sm100_desc.hpp (CUTLASS), which is
// only consulted as documentation -- nothing here includes CUTLASS.
#pragma once
#include <stdlib.h>
#include <cuda.h>
#include "common.cuh"

namespace scn {

// conv_ts.cu: tile-local submanifold kernel; 1 = launched, 0 = not applicable (use conv_tc.cu), < 0 = -status
int conv_ts_try(const float* in, int ld_in, int Cin, const int32_t* map, int n_out, int K, const void* image, const float* bias,
                const float* residual, int ld_res, const float* mask, int ld_mask, float* out, int ld_out, int Cout, int epi,
                float* out2, int ld_out2, int epi2, cudaStream_t stream);
// conv_wgrad_ts.cu: tile-local, deterministic weight gradient; same return convention
int conv_wgrad_ts_try(const float* in, int ld_in, int Cin, const int32_t* map, int n_out, int K, const float* go, int ld_go, int Cout,
                      float* gw, float* gb, cudaStream_t stream);
int make_gather_tmap(CUtensorMap* tm, const float* base, int rows, int C, int ld, CUtensorMapSwizzle swz);

constexpr int TILE_M = 128;
constexpr int KB = 32;                  // fp32 per 128-byte swizzle row
constexpr int A_STAGE_BYTES = TILE_M * 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.  `mbarrier.try_wait` already
// suspends the thread in hardware for a bounded time, so the loop is not a hot spin; roles that wait for a long time
// (epilogue) add a nanosleep back-off.  The wall clock is only consulted every 4096 failed probes: reading
// %globaltimer costs on the order of a microsecond, and taking a start stamp on the FIRST failed probe (as an earlier
// version did) put that microsecond on the critical path of every pipeline unit.
template <int SLEEP_NS = 0>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint64_t t0 = 0;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
        if ((++spins & 4095u) == 0) {
            uint64_t t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t0 == 0) t0 = t1;
            else if (t1 - t0 > 4000000000ull) {
                printf("scn_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
                       threadIdx.x, bar, parity);
                __trap();
            }
        }
    }
}
// One lane of a CONVERGED warp (cute::elect_one_sync).  Single-thread instructions (tcgen05.mma / commit,
// bulk copies) should be issued as `if (elect_one()) ...` from warp-uniform control flow: issued from a divergent
// `if (lane == 0)` region the compiler cannot prove their uniform-register operands warp-uniform and wraps
// every one of them in a vote loop (SASS `BRA.U.ANY`).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst, const void* src, bool valid) {
    int sz = valid ? BYTES : 0;
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    else if constexpr (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// ---- thread-block clusters / distributed shared memory
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// every thread of every CTA of the cluster (release / acquire: shared-memory writes before it are visible after it)
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `addr` (own shared::cta window) in the shared memory of CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ float4 ld_dsmem_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

// all cp.async of this thread, committed to a group or not (wait_group only covers committed groups)
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// arrive on the mbarrier when all cp.async issued so far by this thread have landed (the arrival is one
// of the barrier's expected arrivals: .noinc).  Same mechanism as CUTLASS' sm100 cp.async collective
// (cutlass::arch::cpasync_barrier_arrive_noinc): the producer never blocks on its own copies.
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// the same copy delivered to the same shared-memory offset of every CTA in cta_mask; each destination CTA's barrier (same
// offset) receives the completion bytes
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask)
        : "memory");
}
// arrive (once all tcgen05 operations issued so far by this thread have completed) on the barrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void mma_commit_multicast(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(cta_mask)
                 : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 in
// [0,14), LBO>>4 in [16,30) (unused for swizzled K-major, set to 1), SBO>>4 in [32,46) = 1024 B
// between 8-row groups, version=1 in [46,48), layout_type=2 (SWIZZLE_128B) in [61,64).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor for kind::tf32: c_format=F32 (1<<4), a/b_format=TF32 (2<<7, 2<<10),
// K-major A and B (bits 15,16 = 0), n_dim=N>>3 at [17,23), m_dim=M>>4 at [24,29).
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
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


// MN-major descriptor for 32-bit operands.  CUTLASS: "for mn-major tf32 operands, SW128_32B is the only
// available smem layout" (UMMA::Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> o (1024 bit, 4):(1, 1024 bit)):
// one 128-byte row (32 fp32 of the M/N dimension) per K index, byte-address bits [5,7) ^= bits [7,9),
// i.e. the 32-byte chunk index is XOR-ed with (k_row & 3); K atom = 4 rows = 512 B.
// layout_type = 1 (SWIZZLE_128B_BASE32B); LBO = byte distance between 32-element M/N groups,
// SBO = byte distance between 4-row K groups.
__device__ __forceinline__ uint64_t make_desc_mn_sw128_32b(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (1ull << 61);
}
// position of 16-byte chunk c (0..7) of row r inside a 128-byte row under that swizzle
__device__ __forceinline__ uint32_t swz_mn32b(int r, int c) {
    return (uint32_t)r * 128u + (uint32_t)((((((c >> 1) ^ (r & 3)) << 1) | (c & 1))) << 4);
}
// kind::tf32 instruction descriptor with both operands MN-major (bits 15 and 16 set)
__device__ __forceinline__ uint32_t make_idesc_tf32_mn(int M, int N) {
    return make_idesc_tf32(M, N) | (1u << 15) | (1u << 16);
}

// Row skipping in the gather producers of both tensor-core kernels (conv_tc.cu, conv_wgrad_tc.cu).  SCN_CONV_SKIP=0 / 1
// overrides the default; read per call (a test runs both settings in one process).
#ifndef SCN_CONV_SKIP_DEFAULT
#define SCN_CONV_SKIP_DEFAULT 1
#endif
static inline int conv_row_skipping() {
    const char* e = getenv("SCN_CONV_SKIP");
    if (e && (e[0] == '0' || e[0] == '1')) return e[0] == '1';
    return SCN_CONV_SKIP_DEFAULT;
}

}  // namespace scn
