// Weight gradient of the convolution family on tcgen05 (TF32, fp32 accumulate in TMEM):
//
//   gw[o] (Cin x Cout) = sum_r  in[map[o][r], :]^T (x) go[r, :]
//
// The reduction dimension is the ROW index, so both operands are "MN-major" for the tensor core:
// the gathered tile [128 rows][32 ch] (one 128-byte row per output row, 32-byte chunks XOR-ed with
// row & 3) is the canonical MN-major SW128_32B layout with K = row.  One MMA (K=8) consumes two 4-row
// swizzle atoms; M = 128 input channels (4 channel blocks 16 KB apart), N = up to 128 output channels.
//
// Offset packing: the M = 128 rows of one MMA hold P = 4 / ceil(Cin/32) kernel offsets side by side (4 offsets x 32
// channels for Cin <= 32, 2 x 64 for Cin <= 64, else 1): the four 16 KB channel groups of an A stage are the gathered
// tiles of P different offsets, so one 16-MMA sequence produces the weight gradients of P offsets.  An SS-mode MMA reads
// all 128 M rows (4 KB per K step) from shared memory whether or not they hold channels, and that read is what bounds
// the kernel (profiles/r1_h_issue_bound.md: 101 -> 46 us at C = 32 with the MMAs compiled out).
//
// Work split: grid = (row chunks) x (offset group, 128-wide Cin half, 128-wide Cout half).  A CTA keeps
// the accumulators of ALL its offsets in TMEM across ALL its tiles and adds them to gw with one
// round of atomics at the very end.  The grad-out tile is staged once per tile and reused by every
// offset of the group.  8 producer warps (cp.async gather, zero fill for inactive rows) + 1 MMA warp.
#include <stdlib.h>
#include "tc_common.cuh"

namespace scn {

constexpr int WG_THREADS = 288;

struct WgradParams {
    const float* in;
    int ld_in, Cin;
    const int32_t* map;
    int n_out, K;
    const float* go;
    int ld_go, Cout;
    float* gw;
    float* gb;      // optional: column sums of grad_out (bias gradient) are ADDED here
    int n_tiles, tiles_per_chunk;
    int opg, n_ogroups, n_mhalves;
    int a_stages, g_stages, a_stage_bytes, g_stage_bytes, tmem_cols;
    int skip;      // 1: producers skip rows that are inactive now and were inactive in the stage's previous use
};

template <int VEC>
__device__ __forceinline__ void wg_chunk(uint32_t dst, const float* __restrict__ base, int64_t row_off, bool row_ok,
                                         int col0, int C) {
    constexpr int BYTES = VEC * 4;
#pragma unroll
    for (int j = 0; j < 4 / VEC; ++j) {
        int col = col0 + j * VEC;
        bool valid = row_ok && col < C;
        const float* src = valid ? base + row_off + col : base;
        cp_async<BYTES>(dst + j * BYTES, src, valid);
    }
}

template <int VEC>
__global__ void __launch_bounds__(WG_THREADS, 1) k_conv_wgrad_tc(const WgradParams p) {
    const int chunk = blockIdx.x;
    const int t0 = chunk * p.tiles_per_chunk, t1 = min(p.n_tiles, t0 + p.tiles_per_chunk);
    if (t0 >= t1) return;
    const int og = blockIdx.y % p.n_ogroups;
    const int rest = blockIdx.y / p.n_ogroups;
    const int mh = rest % p.n_mhalves, nh = rest / p.n_mhalves;
    const int o0 = og * p.opg, nO = min(p.K, o0 + p.opg) - o0;
    if (nO <= 0) return;
    const int cin0 = mh * 128, cin_h = min(128, p.Cin - cin0), nblk_a = (cin_h + 31) / 32;
    const int P = nblk_a == 1 ? 4 : (nblk_a == 2 ? 2 : 1);      // offsets packed into one MMA group
    const int bpo = P == 1 ? nblk_a : 4 / P;                     // 16 KB channel blocks per offset
    const int n_mg = (nO + P - 1) / P;                           // MMA groups (= pipeline units) per tile
    const int cout0 = nh * 128, cout_h = min(128, p.Cout - cout0), npad = (cout_h + 15) / 16 * 16;
    const int nblk_g = (npad + 31) / 32;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int AS = p.a_stages, GS = p.g_stages;
    const uint32_t g_base = smem_base + (uint32_t)AS * p.a_stage_bytes;
    const uint32_t bars = g_base + (uint32_t)GS * p.g_stage_bytes;
    auto a_full = [&](int s) { return bars + 8u * s; };
    auto a_empty = [&](int s) { return bars + 8u * (AS + s); };
    auto g_full = [&](int s) { return bars + 8u * (2 * AS + s); };
    auto g_empty = [&](int s) { return bars + 8u * (2 * AS + GS + s); };
    const uint32_t done_bar = bars + 8u * (2 * AS + 2 * GS);
    const uint32_t tmem_slot = done_bar + 8u;

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;      // warp-uniform for the compiler (conv_tc.cu)
    if (tid == 0) {
        for (int s = 0; s < AS; ++s) {
            mbar_init(a_full(s), 64);      // the two owning producer warps
            mbar_init(a_empty(s), 1);
        }
        for (int s = 0; s < GS; ++s) {
            mbar_init(g_full(s), 32);
            mbar_init(g_empty(s), 1);
        }
        mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_trigger();      // PDL (common.cuh): only shared memory / TMEM / parameters were touched so far
    pdl_wait();

    if (warp < 8) {
        // ===================== producers =====================
        // Owned units, as in conv_tc.cu: the CTA's A units (tile, offset) form a flat stream and warp w < NA fills every
        // NA-th of them completely (128 rows x all channel blocks, the only arrivals on that stage's barrier); warps 6 and 7
        // own the grad-out tiles (one per tile, shared by every offset of the group).  The per-unit chain (index read,
        // barrier probe, ring bookkeeping) is paid once per NA units and NA + NG units are in production concurrently.
        // Parity safety needs NA <= AS and NG <= GS (see conv_tc.cu).
        // An A stage is 64 KB, so at most three fit: a PAIR of warps owns a unit (64 rows each, both arrive on the stage
        // barrier) to keep six warps issuing copies -- one warp sustains only ~1 LDGSTS per 60 cycles.
        const int NA = AS < 3 ? AS : 3, NG = GS < 2 ? GS : 2;
        const int c = lane & 7, rsub = lane >> 3;
        const uint32_t dst_lane = swz_mn32b(rsub, c);      // (32 j + rsub + 4 i) & 3 == rsub
        if (warp < 2 * NA) {
            const int own = warp >> 1, half = warp & 1;      // rows 64 * half .. 64 * half + 63 of the unit
            const int n_units = (t1 - t0) * n_mg;
            const uint32_t row_bytes = (uint32_t)p.ld_in * 4u;
            const int cin_lim = min(p.Cin, cin0 + 128);
            const char* in_c = reinterpret_cast<const char*>(p.in + cin0) + c * 16;
            // neighbour indices of the (up to four) offsets of unit u: lane L holds rows 64 half + L and + L + 32
            auto load_idx = [&](int u, int (&dst)[8]) {
#pragma unroll
                for (int q = 0; q < 8; ++q) dst[q] = -1;
                if (u < n_units) {
                    const int tl = u / n_mg, mg = u - tl * n_mg;
                    const int r0 = (t0 + tl) * TILE_M + 64 * half + lane;
#pragma unroll
                    for (int pp = 0; pp < 4; ++pp) {
                        const int o = o0 + mg * P + pp;
                        if (pp < P && o < o0 + nO) {
                            const int32_t* mp = p.map ? p.map + (int64_t)o * p.n_out + r0 : nullptr;
#pragma unroll
                            for (int j = 0; j < 2; ++j)
                                if (r0 + 32 * j < p.n_out) dst[pp * 2 + j] = mp ? __ldg(mp + 32 * j) : r0 + 32 * j;
                        }
                    }
                }
            };
            int idx[8], idx_next[8];
            int sa = own;                                    // own < NA <= AS
            uint32_t pha = 0;
            // Row skipping as in conv_tc.cu: with NA == AS a warp pair always refills the same stage, so bit pp * 2 + j of
            // `dirty` remembers whether rows lane + 32 j of offset slot pp may hold non-zero bytes there (every path below
            // writes all eight chunks of a row, and all channel blocks of a slot see the same rows).  An inactive row over
            // a clean row is not written at all: the fill shares the shared-memory pipe with the MMAs' operand reads.
            const bool skipping = p.skip && NA == AS;
            uint32_t dirty = 0xFFu;                          // shared memory starts undefined
            load_idx(own, idx);
            for (int u = own; u < n_units; u += NA) {
                load_idx(u + NA, idx_next);
                const int mg = u % n_mg;
                mbar_wait(a_empty(sa), pha ^ 1);
                const uint32_t ast = smem_base + (uint32_t)sa * p.a_stage_bytes + dst_lane + (uint32_t)(64 * half) * 128u;
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    if (pp >= P || mg * P + pp >= nO) continue;      // unused offset slot: its accumulator rows are never read
                    int code[2];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int q = pp * 2 + j;
                        code[j] = idx[q] >= 0 ? idx[q] : (((dirty >> q) & 1u) ? -1 : -2);
                        if (skipping) dirty = (dirty & ~(1u << q)) | ((idx[q] >= 0 ? 1u : 0u) << q);
                    }
                    for (int kb = 0; kb < bpo; ++kb) {
                        const int cb = cin0 + kb * KB;               // first channel of this block
                        if (cb >= cin_lim) continue;
                        const uint32_t dst0 = ast + (uint32_t)(pp * bpo + kb) * A_STAGE_BYTES;
                        if (VEC == 4 && cin_lim - cb >= KB) {
                            // one ISETP + IMAD.WIDE + LDGSTS per 16-byte chunk; inactive rows: ignore-src form (zero fill)
                            const char* colp = in_c + kb * (KB * 4);
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
#ifdef SCN_EXP_NOCOPY
                                    if (p.n_out > 0) break;      // timing experiment: no gather copies
#endif
                                    const int r = __shfl_sync(0xffffffffu, code[j], rsub + 4 * i);
                                    const char* src = colp + (uint64_t)(uint32_t)r * row_bytes;
                                    asm volatile(
                                        "{\n\t"
                                        ".reg .pred p, q;\n\t"
                                        "setp.lt.s32 p, %2, 0;\n\t"
                                        "setp.ne.s32 q, %2, -2;\n\t"
                                        "@q cp.async.cg.shared.global [%0], [%1], 16, p;\n\t"
                                        "}" ::"r"(dst0 + (uint32_t)(32 * j + 4 * i) * 128u),
                                        "l"(src), "r"(r)
                                        : "memory");
                                }
                            }
                        } else if (VEC == 4) {
                            // partial channel block, 16-byte copies: the same copy as above behind one lane-constant clamp
                            // (chunks at or beyond the channel limit are zero-filled unless the row is skipped) instead of
                            // per-chunk validity arithmetic (14 -> 6 instructions per copy; conv_tc.cu, profiles/r1_p)
                            const int clamp = cb + c * 4 < cin_lim ? 0x7fffffff : -1;
                            const char* colp = in_c + kb * (KB * 4);
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const int r = min(__shfl_sync(0xffffffffu, code[j], rsub + 4 * i), clamp);
                                    const char* src = colp + (uint64_t)(uint32_t)r * row_bytes;
                                    asm volatile(
                                        "{\n\t"
                                        ".reg .pred p, q;\n\t"
                                        "setp.lt.s32 p, %2, 0;\n\t"
                                        "setp.ne.s32 q, %2, -2;\n\t"
                                        "@q cp.async.cg.shared.global [%0], [%1], 16, p;\n\t"
                                        "}" ::"r"(dst0 + (uint32_t)(32 * j + 4 * i) * 128u),
                                        "l"(src), "r"(r)
                                        : "memory");
                                }
                            }
                        } else {
                            const int col0 = cb + c * 4;
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const int r = __shfl_sync(0xffffffffu, code[j], rsub + 4 * i);
                                    if (r != -2)
                                        wg_chunk<VEC>(dst0 + (uint32_t)(32 * j + 4 * i) * 128u, p.in, (int64_t)r * p.ld_in,
                                                      r >= 0, col0, cin_lim);
                                }
                            }
                        }
                    }
                }
                cp_async_mbar_arrive_noinc(a_full(sa));      // fires when this thread's copies have landed
                sa += NA;
                if (sa >= AS) sa -= AS, pha ^= 1;
#pragma unroll
                for (int q = 0; q < 8; ++q) idx[q] = idx_next[q];
            }
        } else if (warp >= 6 && warp - 6 < NG) {
            // grad-out tiles: dense rows, staged once per tile
            const int gw_i = warp - 6;
            int sg = gw_i;                                   // gw_i < NG <= GS
            uint32_t phg = 0;
            const int cmax = min(p.Cout, cout0 + 128);
            const bool do_bias = p.gb != nullptr && og == 0 && mh == 0;      // one CTA per (row chunk, Cout half)
            float bsum[4] = {0.f, 0.f, 0.f, 0.f};
            for (int tile = t0 + gw_i; tile < t1; tile += NG) {
                mbar_wait(g_empty(sg), phg ^ 1);
                const uint32_t gst = g_base + (uint32_t)sg * p.g_stage_bytes + dst_lane;
                for (int blk = 0; blk < nblk_g; ++blk) {
                    const int col0 = cout0 + blk * KB + c * 4;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = tile * TILE_M + 32 * j + rsub + 4 * i;
                            wg_chunk<VEC>(gst + (uint32_t)blk * A_STAGE_BYTES + (uint32_t)(32 * j + 4 * i) * 128u, p.go,
                                          (int64_t)r * p.ld_go, r < p.n_out, col0, cmax);
                        }
                    }
                }
                cp_async_mbar_arrive_noinc(g_full(sg));
                if (do_bias) {
                    // bias gradient = column sums of grad_out: the tile is in shared memory anyway and this warp has slack
                    // (one tile per n_mg units).  Lane L owns columns L, L+32, ...; the stage is only read by the MMAs.
                    cp_async_wait_all();      // the copies above are not in a committed group: wait_group would not cover them
                    __syncwarp();
                    const uint32_t cc = (uint32_t)lane >> 2, cw = ((uint32_t)lane & 3u) * 4u;
                    const uint32_t gsb = g_base + (uint32_t)sg * p.g_stage_bytes;
#pragma unroll 4
                    for (int r = 0; r < TILE_M; ++r) {
                        const uint32_t off = (uint32_t)r * 128u + (((((cc >> 1) ^ ((uint32_t)r & 3u)) << 1) | (cc & 1u)) << 4) + cw;
#pragma unroll
                        for (int blk = 0; blk < 4; ++blk)
                            if (blk < nblk_g) {
                                float v;
                                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(gsb + (uint32_t)blk * A_STAGE_BYTES + off));
                                bsum[blk] += v;
                            }
                    }
                }
                sg += NG;
                if (sg >= GS) sg -= GS, phg ^= 1;
            }
            if (do_bias) {
#pragma unroll
                for (int blk = 0; blk < 4; ++blk) {
                    const int col = cout0 + blk * KB + lane;
                    if (blk < nblk_g && col < cmax) atomicAdd(p.gb + col, bsum[blk]);
                }
            }
        }
        cp_async_wait_all();
    } else {
        // ===================== MMA issuer =====================
        // Whole warp 8 runs the loop (warp-uniform control flow); one elected lane issues the 16 MMAs of a unit
        // back to back (no per-instruction vote loops).
        const uint32_t idesc = make_idesc_tf32_mn(TILE_M, npad);
        int AS_r, GS_r;
        asm volatile("mov.u32 %0, %1;" : "=r"(AS_r) : "r"(AS));
        asm volatile("mov.u32 %0, %1;" : "=r"(GS_r) : "r"(GS));
        const uint32_t a_bytes = (uint32_t)p.a_stage_bytes, g_bytes = (uint32_t)p.g_stage_bytes;
        int sa = 0, sg = 0;
        uint32_t pha = 0, phg = 0;
        for (int tile = t0; tile < t1; ++tile) {
            mbar_wait(g_full(sg), phg);
            const uint64_t db0 = make_desc_mn_sw128_32b(g_base + (uint32_t)sg * g_bytes, A_STAGE_BYTES, 512);
            for (int mg = 0; mg < n_mg; ++mg) {
                mbar_wait(a_full(sa), pha);
                tc_fence_after();
                const uint64_t da0 = make_desc_mn_sw128_32b(smem_base + (uint32_t)sa * a_bytes, A_STAGE_BYTES, 512);
                const uint32_t tmem_d = tmem_base + (uint32_t)(mg * npad);
                const uint32_t acc0 = tile != t0 ? 1u : 0u;
                if (elect_one()) {
#ifndef SCN_EXP_NOMMA
                    mma_tf32(tmem_d, da0, db0, idesc, acc0);
#endif
#pragma unroll
                    for (int j = 1; j < (
#ifdef SCN_EXP_NOMMA
                                            0 *
#endif
                                            TILE_M / 8);
                         ++j)      // one 8-row K step = 1024 bytes = +64 in the >>4 address field
                        mma_tf32(tmem_d, da0 + (uint64_t)(64 * j), db0 + (uint64_t)(64 * j), idesc, 1u);
                    mma_commit(a_empty(sa));
                }
                __syncwarp();
                if (++sa == AS_r) sa = 0, pha ^= 1;
            }
            if (elect_one()) mma_commit(g_empty(sg));
            __syncwarp();
            if (++sg == GS_r) sg = 0, phg ^= 1;
        }
        if (elect_one()) mma_commit(done_bar);
        __syncwarp();
    }

    if (warp < 4) {
        // ===================== epilogue: TMEM -> atomics into gw =====================
        mbar_wait<500>(done_bar, 0);
        tc_fence_after();
        // accumulator row m = warp * 32 + lane  ->  (packed offset slot pp, channel ci)
        const int pp = warp / bpo;                               // 32 rows per warp = one 16 KB channel block
        const int ci = cin0 + (warp % bpo) * 32 + lane;
        const bool vec_red = (p.Cout & 3) == 0 && (reinterpret_cast<uintptr_t>(p.gw) & 15) == 0;
        for (int mg = 0; mg < n_mg; ++mg) {
            const int o = o0 + mg * P + pp;
            const bool live = pp < P && o < o0 + nO && ci < min(p.Cin, cin0 + 128);
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mg * npad);
            float* dst = p.gw + ((int64_t)o * p.Cin + ci) * p.Cout + cout0;
            for (int c0 = 0; c0 < npad; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + c0, v);
                if (live) {
                    if (vec_red) {      // 16-byte vector reductions: a quarter of the atomic operations
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            if (c0 + j < cout_h)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + j), "f"(v[j]),
                                             "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3])
                                             : "memory");
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < cout_h) atomicAdd(dst + c0 + j, v[j]);
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

}  // namespace scn

using namespace scn;

extern "C" int scn_conv_bwd_weight_fp32(const float* in, int ld_in, int Cin, const int32_t* map, int n_out, int K,
                                        const float* grad_out, int ld_go, int Cout, float* grad_w, scn_stream_t stream);

extern "C" int scn_col_sum_add(const float* in, int ld, int n, int C, float* out, scn_stream_t stream);

extern "C" int scn_conv_bwd_weight(const float* in, int ld_in, int Cin, const int32_t* map, int n_out, int K,
                                   const float* grad_out, int ld_go, int Cout, float* grad_w, float* grad_bias,
                                   int use_tf32, scn_stream_t stream) {
    SCN_REQUIRE(Cin > 0 && Cout > 0 && K > 0, "conv_bwd_weight: bad shape");
    SCN_REQUIRE(map || K == 1, "conv_bwd_weight: identity map requires K == 1");
    if (n_out <= 0) return SCN_OK;
    static int is100 = -1;
    if (is100 < 0) is100 = scn_device_is_sm100();
    if (!use_tf32 || !is100 || Cin > 512 || Cout > 512) {
        int rc = scn_conv_bwd_weight_fp32(in, ld_in, Cin, map, n_out, K, grad_out, ld_go, Cout, grad_w, stream);
        if (rc || !grad_bias) return rc;
        return scn_col_sum_add(grad_out, ld_go, n_out, Cout, grad_bias, stream);
    }

    {      // Morton-ordered level with a tile book, C -> C, 3^3: the tile-local deterministic kernel (conv_wgrad_ts.cu)
        const int rc = scn::conv_wgrad_ts_try(in, ld_in, Cin, map, n_out, K, grad_out, ld_go, Cout, grad_w, grad_bias, as_stream(stream));
        if (rc < 0) return -rc;
        if (rc == 1) return SCN_OK;
    }
    WgradParams p;
    p.in = in, p.ld_in = ld_in, p.Cin = Cin, p.map = map, p.n_out = n_out, p.K = K;
    p.go = grad_out, p.ld_go = ld_go, p.Cout = Cout, p.gw = grad_w, p.gb = grad_bias;
    p.n_tiles = cdiv(n_out, TILE_M);
    p.skip = scn::conv_row_skipping();
    const int cin_h = Cin < 128 ? Cin : 128, cout_h = Cout < 128 ? Cout : 128;
    const int npad = (cout_h + 15) / 16 * 16;
    p.n_mhalves = cdiv(Cin, 128);
    const int n_nhalves = cdiv(Cout, 128);
    p.a_stage_bytes = cdiv(cin_h, 32) * A_STAGE_BYTES;
    p.g_stage_bytes = cdiv(npad, 32) * A_STAGE_BYTES;
    // offset packing: P offsets share one MMA group (see the header); a CTA keeps ceil(opg / P) groups of npad TMEM
    // columns each, at most 512 columns
    const int nblk_a = cdiv(cin_h, 32);
    const int P = nblk_a == 1 ? 4 : (nblk_a == 2 ? 2 : 1);
    int max_mg = 512 / npad;
    if (max_mg < 1) max_mg = 1;
    p.opg = max_mg * P;
    if (p.opg > K) p.opg = K;
    p.n_ogroups = cdiv(K, p.opg);
    p.opg = cdiv(cdiv(K, p.n_ogroups), P) * P;          // balance the groups, keep them multiples of P
    p.n_ogroups = cdiv(K, p.opg);
    int tc = 32;
    while (tc < cdiv(p.opg, P) * npad) tc <<= 1;
    p.tmem_cols = tc;
    // shared memory, one CTA per SM: an A stage is always the full 64 KB window an M = 128 MMA reads (four 16 KB channel
    // blocks = P offsets x Cin/P channels); one or two grad-out stages; as many A stages as fit (one owner warp each)
    p.a_stage_bytes = 4 * A_STAGE_BYTES;
    p.g_stages = (2 * p.a_stage_bytes + 2 * p.g_stage_bytes <= 222 * 1024) ? 2 : 1;
    p.a_stages = (222 * 1024 - p.g_stages * p.g_stage_bytes) / p.a_stage_bytes;
    if (p.a_stages > 6) p.a_stages = 6;
    SCN_REQUIRE(p.a_stages >= 2, "conv_bwd_weight: tile does not fit in shared memory (Cin=%d Cout=%d)", Cin, Cout);
    int smem = p.a_stages * p.a_stage_bytes + p.g_stages * p.g_stage_bytes;
    const int ctas_per_sm = 1;
    smem += 1024 + 256;
    // every CTA ends with opg x Cin x Cout reductions into gw: keep at least 4 tiles per chunk so that this does not
    // dominate small levels, and when that leaves fewer chunks than SMs split the offsets over more groups instead
    const int max_chunks = p.n_tiles >= 8 ? p.n_tiles / 4 : (p.n_tiles >= 2 ? 2 : 1);
    {
        const int halves = p.n_mhalves * n_nhalves;
        int want = cdiv(sm_count(), max_chunks * halves);
        const int most = cdiv(K, P);
        if (want > most) want = most;
        if (want > p.n_ogroups) {
            p.opg = cdiv(cdiv(K, want), P) * P;
            p.n_ogroups = cdiv(K, p.opg);
            tc = 32;
            while (tc < cdiv(p.opg, P) * npad) tc <<= 1;
            p.tmem_cols = tc;
        }
    }
    const int groups_y = p.n_ogroups * p.n_mhalves * n_nhalves;
    int n_chunks = (sm_count() * ctas_per_sm + groups_y - 1) / groups_y;
    if (n_chunks > max_chunks) n_chunks = max_chunks;
    if (n_chunks < 1) n_chunks = 1;
    p.tiles_per_chunk = cdiv(p.n_tiles, n_chunks);
    n_chunks = cdiv(p.n_tiles, p.tiles_per_chunk);
    int vec = 1;
    const uintptr_t al = reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(grad_out);
    if (Cin % 4 == 0 && Cout % 4 == 0 && ld_in % 4 == 0 && ld_go % 4 == 0 && (al & 15) == 0) vec = 4;
    else if (Cin % 2 == 0 && Cout % 2 == 0 && ld_in % 2 == 0 && ld_go % 2 == 0 && (al & 7) == 0) vec = 2;
    dim3 grid(n_chunks, groups_y);
    cudaError_t e = cudaSuccess;
    auto launch = [&](auto kern) {
        e = (cudaError_t)scn::ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem);
        if (e != cudaSuccess) return;
        scn::PdlLaunch L(grid, dim3(WG_THREADS), smem, as_stream(stream));
        e = cudaLaunchKernelEx(&L.cfg, kern, p);
    };
    if (vec == 4) launch(k_conv_wgrad_tc<4>);
    else if (vec == 2) launch(k_conv_wgrad_tc<2>);
    else launch(k_conv_wgrad_tc<1>);
    if (e != cudaSuccess) {
        cudaGetLastError();
        scn::set_error("conv_bwd_weight: cudaFuncSetAttribute(%d bytes): %s", smem, cudaGetErrorString(e));
        return SCN_ERR_CUDA;
    }
    return check_launch("conv_bwd_weight_tc");
}
