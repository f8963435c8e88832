// Shared helpers for the scn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/scn_b200.h"

namespace scn {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define SCN_REQUIRE(cond, ...)                      \
    do {                                            \
        if (!(cond)) {                              \
            scn::set_error(__VA_ARGS__);            \
            return SCN_ERR_INVALID;                 \
        }                                           \
    } while (0)

static inline cudaStream_t as_stream(scn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// grid size for grid-stride kernels: a multiple of the SM count (148 on B200), capped by work
int sm_count();
// cudaFuncAttributeMaxDynamicSharedMemorySize only has to grow: remember the largest value set per kernel (a driver
// call per launch costs microseconds of host time and the small levels are host bound).  Returns a cudaError_t.
int ensure_dynamic_smem(const void* kernel, int bytes);
static inline int grid_for(int64_t work_items, int block, int ctas_per_sm = 8) {
    int64_t need = (work_items + block - 1) / block;
    int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Kernels that are launched back to back on one stream inside a C-ABI call (a residual unit = 3 launches forward, 6
// backward) pay ~3 us of launch latency + prologue each, serialised behind the previous kernel's drain.  With the
// programmatic-stream-serialization launch attribute the next kernel's CTAs may start (barrier init, TMEM allocation,
// index arithmetic) as soon as every CTA of the previous kernel has executed `pdl_trigger()` or exited and resources are
// free; `pdl_wait()` then blocks until the previous grid has COMPLETED and its memory is visible.  Rule kept by every
// kernel that uses it: no global-memory access before pdl_wait().
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// SCN_PDL=0 turns the launch attribute off (the device-side instructions are then no-ops)
bool pdl_enabled();
// Per-thread exclusion list: launches on these streams go without the attribute while the scope lives.  With TWO streams
// feeding the GPU (scn_unet_bwd: input-gradient chain + weight gradients) an early-launched dependent grid parks its CTAs
// -- shared memory and tensor memory included -- in griddepcontrol.wait on SMs the OTHER stream's ready kernel could use
// (measured: backward call 4.17 ms with the attribute on both streams, 3.91 ms without).
struct PdlMask {
    int n = 0;
    cudaStream_t s[6];
};
PdlMask& pdl_mask();
static inline bool pdl_allowed(cudaStream_t st) {
    const PdlMask& m = pdl_mask();
    for (int i = 0; i < m.n; ++i)
        if (m.s[i] == st) return false;
    return true;
}
struct PdlMaskScope {
    PdlMask saved;
    PdlMaskScope() : saved(pdl_mask()) {}
    void exclude(cudaStream_t st) {
        PdlMask& m = pdl_mask();
        if (m.n < 6) m.s[m.n++] = st;
    }
    ~PdlMaskScope() { pdl_mask() = saved; }
};
// launch config with the PDL attribute (and an optional cluster dimension); attrs must outlive the launch call
struct PdlLaunch {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attrs[2];
    PdlLaunch(dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster = 0) {
        cfg = cudaLaunchConfig_t{};
        cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
        int n = 0;
        if (pdl_enabled() && pdl_allowed(st)) {
            attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attrs[n].val.programmaticStreamSerializationAllowed = 1;
            ++n;
        }
        if (cluster > 1) {
            attrs[n].id = cudaLaunchAttributeClusterDimension;
            attrs[n].val.clusterDim.x = cluster, attrs[n].val.clusterDim.y = 1, attrs[n].val.clusterDim.z = 1;
            ++n;
        }
        cfg.attrs = attrs, cfg.numAttrs = n;
    }
};

// ---------------------------------------------------------------- second output of a convolution epilogue
// out2 = epi2(out): ReLU and / or round-to-nearest TF32 of the value just written -- the operand the NEXT gather-GEMM reads
// (the residual chain keeps the unrounded value), saving that layer's separate elementwise pass.
__device__ __forceinline__ float epi2_apply(float x, int epi2) {
    if (epi2 & SCN_EPI_RELU) x = fmaxf(x, 0.f);
    if (epi2 & SCN_EPI_ROUND) {
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x));
        x = __uint_as_float(t);
    }
    return x;
}
__device__ __forceinline__ float4 epi2_apply4(float4 v, int epi2) {
    return make_float4(epi2_apply(v.x, epi2), epi2_apply(v.y, epi2), epi2_apply(v.z, epi2), epi2_apply(v.w, epi2));
}

// ---------------------------------------------------------------- packed voxel keys
__host__ __device__ __forceinline__ uint64_t make_key(uint32_t x, uint32_t y, uint32_t z, uint32_t b) {
    return ((uint64_t)b << 48) | ((uint64_t)x << 32) | ((uint64_t)y << 16) | (uint64_t)z;
}
__host__ __device__ __forceinline__ int key_x(uint64_t k) { return (int)((k >> 32) & 0xFFFF); }
__host__ __device__ __forceinline__ int key_y(uint64_t k) { return (int)((k >> 16) & 0xFFFF); }
__host__ __device__ __forceinline__ int key_z(uint64_t k) { return (int)(k & 0xFFFF); }
__host__ __device__ __forceinline__ int key_b(uint64_t k) { return (int)((k >> 48) & 0xFFFF); }

// ---------------------------------------------------------------- open-addressing hash
__device__ __forceinline__ uint32_t hash_key(uint64_t k) {
    // murmur3 fmix64
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return (uint32_t)k;
}

// returns slot of `key` (inserting it if absent)
__device__ __forceinline__ uint32_t hash_insert_slot(uint64_t* tab_keys, uint32_t mask, uint64_t key) {
    uint32_t s = hash_key(key) & mask;
    while (true) {
        unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(tab_keys + s),
                                            (unsigned long long)SCN_EMPTY_KEY, (unsigned long long)key);
        if (prev == SCN_EMPTY_KEY || prev == key) return s;
        s = (s + 1) & mask;
    }
}

// returns slot of key or 0xFFFFFFFF
__device__ __forceinline__ uint32_t hash_find_slot(const uint64_t* __restrict__ tab_keys, uint32_t mask,
                                                   uint64_t key) {
    uint32_t s = hash_key(key) & mask;
    while (true) {
        uint64_t k = __ldg(tab_keys + s);
        if (k == key) return s;
        if (k == SCN_EMPTY_KEY) return 0xFFFFFFFFu;
        s = (s + 1) & mask;
    }
}

__device__ __forceinline__ int hash_lookup(const uint64_t* __restrict__ tab_keys,
                                           const int32_t* __restrict__ tab_vals, uint32_t mask,
                                           uint64_t key) {
    uint32_t s = hash_find_slot(tab_keys, mask, key);
    return s == 0xFFFFFFFFu ? -1 : __ldg(tab_vals + s);
}

}  // namespace scn
