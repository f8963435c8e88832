// Probe for TMA tile::gather4 semantics on sm_100a: which boxDim the tensor map needs, how rows land in
// shared memory under SWIZZLE_128B, and whether a negative row index zero-fills.
// usage: gather4_probe <box_rows> <dtype 0=f32 1=tf32>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int r0, int r1, int r2, int r3, int col) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    uint32_t sbase = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
    uint32_t sbar = (uint32_t)__cvta_generic_to_shared(&bar);
    float* sm = (float*)(smem + (sbase - (uint32_t)__cvta_generic_to_shared(smem)));
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = -7.f;     // 4 KB sentinel
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sbar), "r"(4 * 128) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
            ::"r"(sbase + 512u), "l"(&tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(sbar)
            : "memory");
        uint32_t ok = 0;
        long spins = 0;
        while (!ok && spins < 20000000) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(sbar) : "memory");
            ++spins;
        }
        out[1024] = ok ? 1.f : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = sm[i];
}

int main(int argc, char** argv) {
    int box_rows = argc > 1 ? atoi(argv[1]) : 1;
    int dtype = argc > 2 ? atoi(argv[2]) : 0;
    int swz = argc > 3 ? atoi(argv[3]) : 0;   // 0 = SWIZZLE_128B, 1 = SWIZZLE_128B_ATOM_32B
    int Ccols = argc > 4 ? atoi(argv[4]) : 32; int col0 = argc > 5 ? atoi(argv[5]) : 0;
    const int R = 64; const int C = Ccols;
    float* h = (float*)malloc(R * C * 4);
    for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = r * 100 + c + 0.0001234f;
    // rounding probes (row 9): 1 + 0.75 ulp_tf32 -> RN gives 1+2^-10, truncation gives 1; 1 + 0.5 ulp (tie); 1 + 0.25 ulp
    h[9 * C + 0] = 1.0f + 0.75f / 1024.f; h[9 * C + 4] = 1.0f + 0.5f / 1024.f; h[9 * C + 8] = 1.0f + 0.25f / 1024.f;
    h[9 * C + 12] = -(1.0f + 0.75f / 1024.f);
    float *d, *out;
    cudaMalloc(&d, R * C * 4);
    cudaMalloc(&out, 1025 * 4);
    cudaMemcpy(d, h, R * C * 4, cudaMemcpyHostToDevice);
    EncodeTiled enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q) != cudaSuccess || !enc) {
        printf("no cuTensorMapEncodeTiled\n");
        return 2;
    }
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)R};
    cuuint64_t gstride[1] = {(cuuint64_t)C * 4};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = enc(&tmap, dtype ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swz ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d box_rows=%d dtype=%d\n", (int)rc, box_rows, dtype);
    if (rc) return 3;
    probe<<<1, 128, 8192>>>(tmap, out, 5, -1, 2, 9, col0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e) return 4;
    float res[1025];
    cudaMemcpy(res, out, sizeof(res), cudaMemcpyDeviceToHost);
    printf("barrier completed: %g\n", res[1024]);
    // rows of 128 B (32 floats): print the first float of each 16-byte chunk for the 8 rows after offset 512 B
    for (int row = 4; row < 8; ++row) {
        printf("smem row %2d (byte %4d):", row, row * 128);
        for (int ch = 0; ch < 8; ++ch) printf(" %11.6f", res[row * 32 + ch * 4]);
        printf("\n");
    }
    return 0;
}
