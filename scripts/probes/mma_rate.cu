// Probe: how fast can tcgen05.mma kind::tf32 (M = 128, K = 8) be issued?  One or two issuing warps per CTA, SS and TS
// form, N in {32, 64, 128}.  Prints cycles per MMA instruction (clock64 around the issue loop incl. the final commit wait).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu -I ../../sparse_rcnn_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace scn;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                 "r"(a), "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}

template <int N, bool TS>
__global__ void __launch_bounds__(128) k_probe(int n_units, int issuers, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_smem = base, b_smem = base + 16384, bars = base + 16384 + N * 128, slot = bars + 64;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (16384 + N * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
    if (tid == 0) {
        mbar_init(bars, 1), mbar_init(bars + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    if (warp < issuers) {
        const uint32_t idesc = make_idesc_tf32(128, N);
        const uint64_t da = make_desc_sw128(a_smem), db = make_desc_sw128(b_smem);
        const uint32_t d = tmem + warp * 256, a_t = tmem + 128 + warp * 256;
        const long long t0 = clock64();
        for (int u = 0; u < n_units; ++u) {
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (TS) mma_ts(d, a_t + k * 8, db + 2 * k, idesc, 1u);
                    else mma_tf32(d, da + 2 * k, db + 2 * k, idesc, 1u);
                }
            }
            __syncwarp();
        }
        const long long t1 = clock64();
        if (elect_one()) mma_commit(bars + 8 * warp);
        mbar_wait(bars + 8 * warp, 0);
        const long long t2 = clock64();
        if ((tid & 31) == 0 && blockIdx.x == 0) out[2 * warp] = t1 - t0, out[2 * warp + 1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int N, bool TS>
void run(int issuers) {
    long long* d;
    cudaMalloc(&d, 64);
    const int units = 2000, smem = 1024 + 16384 + N * 128 + 256;
    cudaFuncSetAttribute(k_probe<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int r = 0; r < 2; ++r) k_probe<N, TS><<<148, 128, smem>>>(units, issuers, d);
    long long h[4] = {0, 0, 0, 0};
    cudaError_t e = cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("N=%3d %s issuers=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (per issuing warp)%s  [%s]\n", N, TS ? "TS" : "SS", issuers,
           h[0] / (4.0 * units), h[1] / (4.0 * units), issuers > 1 ? "" : "", cudaGetErrorString(e));
    if (issuers > 1) printf("       second warp: issue %.1f, complete %.1f\n", h[2] / (4.0 * units), h[3] / (4.0 * units));
    cudaFree(d);
}

int main() {
    run<32, false>(1), run<32, true>(1), run<32, false>(2), run<32, true>(2);
    run<64, true>(1), run<64, true>(2), run<128, true>(1), run<48, true>(1), run<16, true>(1);
    return 0;
}
