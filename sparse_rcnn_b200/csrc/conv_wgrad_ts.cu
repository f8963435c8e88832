// Tile-local weight (+ bias) gradient of submanifold 3^3 convolutions on Morton-ordered rows, TF32 tcgen05, deterministic.
//
//   gw[o][ci][co] += sum_r  in[map[o][r]][ci] * go[r][co]            gb[co] += sum_r go[r][co]
//
// conv_wgrad_tc.cu gathers the 128 input rows of every (tile, offset) pair from L2 with cp.async into a 16 KB shared-memory
// operand (zero fill for the ~60 % missing neighbours) and adds its accumulators to gw with floating-point atomics.  Here,
// as in conv_ts.cu, a tile's distinct input rows (its halo set, ~280 rows) are copied into shared memory ONCE, and the
// A operand is built in TENSOR MEMORY:
//   * the reduction runs over rows, so A is [M = (offset, input channel)] x [K = row]: TMEM lane = (offset slot, channel),
//     TMEM column = row of the tile.  A gather warp owns 32 lanes = 32 channels of ONE offset; for row r it reads the
//     128 contiguous bytes of the neighbour's halo row (one conflict-free LDS.32 wavefront) -- and because all its lanes
//     share (offset, row), a missing neighbour is a warp-uniform skip: only ACTIVE pairs cost a shared-memory access
//     (1274 of 3456 per tile at level 0), the neighbour codes are read eight at a time by broadcast;
//   * B = the tile's grad-out rows in shared memory (MN-major SW128_32B, staged once per tile, reused by all offsets);
//   * tcgen05.mma kind::tf32, A from TMEM (K-major), M = 128 lanes = P offsets x C channels side by side, N = C, K = 8
//     rows per instruction; the fp32 accumulators of ALL offsets of a CTA stay in TMEM across all its tiles;
//   * at the end every CTA writes its accumulators to its own slice of a workspace and a second kernel adds the slices to
//     gw in CTA order: no floating-point atomics, the gradient is bit-reproducible (conv_wgrad_tc.cu's is not).
// One persistent CTA per SM (C = 48: two CTAs per tile stream, 14 / 13 offsets each, because 27 x 48 accumulator lanes
// x 48 columns exceed the 512 TMEM columns).  Warps: 4 G gather (G groups x 4 lane quarters, one 64-row unit each) |
// 2 halo loaders | grad-out loader (+ bias column sums) | MMA issuer; group 0 also drains the accumulators at the end.
#include <atomic>
#include <mutex>
#include <unordered_map>
#include "tile_book.cuh"

namespace scn {

struct WgradTsParams {
    const float* in;
    int ld_in;
    TileBook book;
    const int32_t* map;      // [27][n_out]: rows whose halo index exceeds the shared-memory buffer
    const float* go;
    int ld_go;
    float* part;             // [grid][NMG * 128][C] accumulators of every CTA
    float* part_b;           // [ctas_per_group][C] bias column sums (offset group 0 only) or nullptr
    int cap, ctas_per_group;
};

template <int C>
struct WtShape {
    static constexpr int W = C <= 32 ? 32 : 64;                   // lanes per offset slot
    static constexpr int P = 128 / W;                             // offsets side by side in one M = 128 group
    static constexpr int OG = C <= 32 ? 1 : 2;                    // CTAs sharing the 27 offsets of a tile stream
    static constexpr int OFFS = (TS_K + OG - 1) / OG;             // offsets per CTA (the last group may hold fewer)
    static constexpr int NMG = (OFFS + P - 1) / P;                // M groups = accumulators per CTA
    static constexpr int DCOLS = NMG * C;
    static constexpr int NS = (512 - DCOLS) / 64;                 // A stages of 64 columns (= 64 rows of the tile)
    static constexpr int G = NS < 4 ? NS : 4;
    static constexpr int NU = NMG * 2;                            // units per tile: (M group, row half)
    static constexpr int NBLK = (C + 31) / 32;
    static constexpr int GO_BYTES = NBLK * A_STAGE_BYTES;
    static constexpr int PITCH = C * 4 + 16;
    static constexpr int CPR = C / 4;
    static constexpr int THREADS = (4 * G + 4) * 32;
    static constexpr int N_BARS = 8 + 2 * NS + 1;
    static constexpr int FIXED_SMEM = 1024 + 2 * GO_BYTES + 2 * TS_BLOB_BYTES + 512;
    static_assert(NS >= 2 && G >= 2 && NU >= G, "tile-local weight gradient: shape does not fit tensor memory");
};

__device__ __forceinline__ void mma_tf32_ts_mn(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}

template <int C>
__global__ void __launch_bounds__(WtShape<C>::THREADS, 1) k_wgrad_ts(const WgradTsParams p) {
    using S = WtShape<C>;
    constexpr int NS = S::NS, G = S::G, NU = S::NU, PITCH = S::PITCH, CPR = S::CPR, NMG = S::NMG;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t go0 = smem_base;                                       // swizzled operand: 1024-byte aligned
    const uint32_t halo_bytes = (uint32_t)p.cap * PITCH;
    const uint32_t halo0 = go0 + 2u * S::GO_BYTES;
    const uint32_t blob0 = halo0 + 2u * halo_bytes;
    const uint32_t bars = blob0 + 2u * TS_BLOB_BYTES;
    auto halo_full = [&](int b) { return bars + 8u * b; };
    auto halo_empty = [&](int b) { return bars + 8u * (2 + b); };
    auto go_full = [&](int b) { return bars + 8u * (4 + b); };
    auto go_empty = [&](int b) { return bars + 8u * (6 + b); };
    auto a_full = [&](int s) { return bars + 8u * (8 + s); };
    auto a_empty = [&](int s) { return bars + 8u * (8 + NS + s); };
    const uint32_t done_bar = bars + 8u * (8 + 2 * NS);
    const uint32_t tmem_slot = bars + 8u * S::N_BARS;

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    constexpr int W_LOAD0 = 4 * G, W_GO = W_LOAD0 + 2, W_MMA = W_LOAD0 + 3;
    // this CTA: offset group og, tiles cta_in_group, cta_in_group + ctas_per_group, ...
    const int og = blockIdx.x / p.ctas_per_group, cig = blockIdx.x - og * p.ctas_per_group;
    const int o_first = og * S::OFFS;
    const int n_off = (TS_K - o_first) < S::OFFS ? (TS_K - o_first) : S::OFFS;
    const int n_tiles = p.book.n_tiles, stride = p.ctas_per_group;
    const int my_tiles = cig < n_tiles ? (n_tiles - cig + stride - 1) / stride : 0;

    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(halo_full(b), 64 + 1);      // two loader warps (cp.async arrivals) + the blob's bulk copy
            mbar_init(halo_empty(b), 4 * G);      // every gather warp
            mbar_init(go_full(b), 32);
            mbar_init(go_empty(b), 1);
        }
        for (int s = 0; s < NS; ++s) {
            mbar_init(a_full(s), 4);              // the four lane quarters of the unit
            mbar_init(a_empty(s), 1);
        }
        mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_trigger();
    pdl_wait();
    constexpr uint32_t A_COL0 = (uint32_t)S::DCOLS;

    if (warp < W_LOAD0) {
        // ===================== gather warps: halo rows -> registers -> tensor memory (lane = channel, column = row) ==========
        const int g = warp >> 2, q = warp & 3;
        const int pp = (32 * q) / S::W, ch = (32 * q) % S::W + lane;      // offset slot inside the M group, input channel
        const bool ch_ok = ch < C;
        const uint32_t cap = (uint32_t)p.cap;
        const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + A_COL0;
        int cur_it = 0;                                  // tile iteration this warp is working on
        uint32_t seq = 0;
        bool have_tile = false, tile_far = false;
#pragma unroll 1
        for (int cnt = g; cnt < my_tiles * NU; cnt += G) {
            const int it = cnt / NU, u = cnt - it * NU;
            const int mg = u >> 1, hf = u & 1;
            if (!have_tile || it != cur_it) {
                if (have_tile) {      // done with the previous tile's halo
                    __syncwarp();
                    if (elect_one()) mbar_arrive(halo_empty(cur_it & 1));
                }
                cur_it = it, have_tile = true;
                seq = __ldg(p.book.useq + cig + it * stride);
                {
                    const int nl = __ldg(p.book.nloc + cig + it * stride);      // rows stored (clipped to TS_ROWS_CAP)
                    tile_far = nl > (int)cap || nl >= TS_ROWS_CAP;
                }
                mbar_wait(halo_full(it & 1), (uint32_t)(it >> 1) & 1u);
            }
            const int buf = it & 1, tile = cig + it * stride;
            const uint32_t hb = halo0 + (uint32_t)buf * halo_bytes + (uint32_t)ch * 4u;
            const uint32_t blob = blob0 + (uint32_t)buf * TS_BLOB_BYTES;
            const int ol = mg * S::P + pp, o = o_first + ol;
            const bool live = ol < n_off && ((seq >> seq_bit((uint32_t)o)) & 1u);      // the offset has active pairs in this tile
            const int stage = cnt % NS;
            const uint32_t par = (uint32_t)(cnt / NS) & 1u;
            uint32_t v[32];
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int r0 = hf * 64 + half * 32;      // first row (= column) of this pass
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
                if (live) {
                    // 32 neighbour codes of this pass: four broadcast loads (the same address in every lane), issued together
                    const uint32_t ca = blob + TS_BLOB_LMAP + (uint32_t)(o * TILE_M + r0) * 2u;
                    uint32_t c4[16];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(c4[4 * jj]), "=r"(c4[4 * jj + 1]), "=r"(c4[4 * jj + 2]), "=r"(c4[4 * jj + 3])
                                     : "r"(ca + 16u * jj));
                    // any row of this pass outside the shared-memory halo buffer (index >= cap, not "no neighbour")?  rare
                    const bool far = tile_far;      // only tiles whose halo list exceeds the buffer can hold such rows
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const uint32_t code = (c4[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
                        // branch-free: all lanes share (offset, row), a missing neighbour predicates the load off
                        asm volatile(
                            "{\n\t"
                            ".reg .pred p;\n\t"
                            "setp.lt.u32 p, %1, %2;\n\t"
                            "@p ld.shared.b32 %0, [%3];\n\t"
                            "}"
                            : "+r"(v[j])
                            : "r"(ch_ok ? code : 0xFFFFu), "r"(cap), "r"(hb + code * PITCH));
                    }
                    if (far) {
#pragma unroll 1
                        for (int j = 0; j < 32; ++j) {
                            uint32_t code;      // re-read: indexing c4 with a loop variable would put it in local memory
                            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(code) : "r"(ca + 2u * j));
                            if (code >= cap && code != TS_INACTIVE) {
                                const int src = __ldg(p.map + (int64_t)o * p.book.n_out + tile * TILE_M + r0 + j);
                                const uint32_t x = ch_ok ? __float_as_uint(__ldg(p.in + (int64_t)src * p.ld_in + ch)) : 0u;
#pragma unroll
                                for (int k = 0; k < 32; ++k)
                                    if (k == j) v[k] = x;
                            }
                        }
                    }
                }
                if (half == 0) {
                    mbar_wait(a_empty(stage), par ^ 1u);      // the MMAs that read this stage's previous unit are done
                    tc_fence_after();
                }
                tmem_st32(tq + (uint32_t)(stage * 64 + half * 32), v);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive(a_full(stage));
        }
        if (have_tile) {
            __syncwarp();
            if (elect_one()) mbar_arrive(halo_empty(cur_it & 1));
        }
        if (g == 0 && my_tiles > 0) {
            // ===================== drain: accumulators -> this CTA's slice of the workspace =====================
            mbar_wait<500>(done_bar, 0);
            tc_fence_after();
            float* dst = p.part + ((int64_t)blockIdx.x * NMG * 128 + q * 32 + lane) * C;
#pragma unroll 1
            for (int mg = 0; mg < NMG; ++mg) {
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mg * C);
#pragma unroll 1
                for (int c0 = 0; c0 < C; c0 += 16) {
                    float a[16];
                    tmem_ld16(taddr + c0, a);
                    float4* d4 = reinterpret_cast<float4*>(dst + (int64_t)mg * 128 * C + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) d4[j] = make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
                }
            }
            tc_fence_before();
        }
    } else if (warp == W_LOAD0 || warp == W_LOAD0 + 1) {
        // ===================== halo loaders (as conv_ts.cu): a tile's distinct input rows, once =====================
        const int lw = warp - W_LOAD0;
        const char* in_c = reinterpret_cast<const char*>(p.in);
        const uint32_t row_bytes = (uint32_t)p.ld_in * 4u;
        int cnt_next = my_tiles > 0 ? __ldg(p.book.nloc + cig) : 0;
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = cig + it * stride, buf = it & 1;
            int cnt = cnt_next;
            if (it + 1 < my_tiles) cnt_next = __ldg(p.book.nloc + tile + stride);
            if (cnt > p.cap) cnt = p.cap;
            const int32_t* rl = p.book.rows + (int64_t)tile * TS_ROWS_CAP;
            constexpr int ROUNDS = TS_ROWS_CAP / 64;
            int idx[ROUNDS];
#pragma unroll
            for (int k = 0; k < ROUNDS; ++k) {
                const int at = lw * 32 + 64 * k + lane;
                idx[k] = at < cnt ? __ldg(rl + at) : -1;
            }
            mbar_wait(halo_empty(buf), ((uint32_t)(it >> 1) & 1u) ^ 1u);
            const uint32_t hb = halo0 + (uint32_t)buf * halo_bytes;
            if (lw == 0 && elect_one()) {
                mbar_arrive_expect_tx(halo_full(buf), TS_BLOB_BYTES);
                bulk_g2s(blob0 + (uint32_t)buf * TS_BLOB_BYTES, p.book.blobs + (int64_t)tile * TS_BLOB_BYTES, TS_BLOB_BYTES,
                         halo_full(buf));
            }
#pragma unroll
            for (int k = 0; k < ROUNDS; ++k) {
                const int base = lw * 32 + 64 * k;
                if (base < cnt) {
                    const int mine = idx[k];
#pragma unroll
                    for (int i = 0; i < CPR; ++i) {
                        const int t = lane + 32 * i;
                        const int jr = t / CPR, c = t % CPR;
                        const int ridx = __shfl_sync(0xffffffffu, mine, jr);
                        if (ridx >= 0) {
                            const char* src = in_c + (uint64_t)(uint32_t)ridx * row_bytes + c * 16;
                            const uint32_t dst = hb + (uint32_t)(base + jr) * PITCH + (uint32_t)(c * 16);
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                        }
                    }
                }
            }
            cp_async_mbar_arrive_noinc(halo_full(buf));
        }
        cp_async_wait_all();
    } else if (warp == W_GO) {
        // ===================== grad-out tiles (MN-major SW128_32B, one per tile) + bias column sums =====================
        const int c8 = lane & 7, rsub = lane >> 3;
        const uint32_t dst_lane = swz_mn32b(rsub, c8);
        const bool do_bias = p.part_b != nullptr && og == 0;
        float bsum[S::NBLK];
#pragma unroll
        for (int b = 0; b < S::NBLK; ++b) bsum[b] = 0.f;
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = cig + it * stride, buf = it & 1;
            mbar_wait(go_empty(buf), ((uint32_t)(it >> 1) & 1u) ^ 1u);
            const uint32_t gst = go0 + (uint32_t)buf * S::GO_BYTES;
#pragma unroll
            for (int blk = 0; blk < S::NBLK; ++blk) {
                const int col = blk * KB + c8 * 4;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = tile * TILE_M + 32 * j + rsub + 4 * i;
                        const bool valid = r < p.book.n_out && col < C;
                        const float* src = valid ? p.go + (int64_t)r * p.ld_go + col : p.go;
                        cp_async<16>(gst + (uint32_t)blk * A_STAGE_BYTES + (uint32_t)(32 * j + 4 * i) * 128u + dst_lane, src, valid);
                    }
                }
            }
            cp_async_mbar_arrive_noinc(go_full(buf));
            if (do_bias) {
                // column sums of the tile from shared memory (this warp has slack); lane L owns column L of every 32-column block
                cp_async_wait_all();
                __syncwarp();
                const uint32_t cc = (uint32_t)lane >> 2, cw = ((uint32_t)lane & 3u) * 4u;
#pragma unroll 4
                for (int r = 0; r < TILE_M; ++r) {
                    const uint32_t off = (uint32_t)r * 128u + (((((cc >> 1) ^ ((uint32_t)r & 3u)) << 1) | (cc & 1u)) << 4) + cw;
#pragma unroll
                    for (int blk = 0; blk < S::NBLK; ++blk) {
                        float x;
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(gst + (uint32_t)blk * A_STAGE_BYTES + off));
                        bsum[blk] += x;
                    }
                }
            }
        }
        cp_async_wait_all();
        if (do_bias && my_tiles > 0) {
#pragma unroll
            for (int blk = 0; blk < S::NBLK; ++blk) {
                const int col = blk * KB + lane;
                if (col < C) p.part_b[(int64_t)cig * C + col] = bsum[blk];
            }
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer: A from tensor memory (K-major), B = grad-out tile (MN-major) =====================
        const uint32_t idesc = make_idesc_tf32(TILE_M, C) | (1u << 16);      // B MN-major
#pragma unroll 1
        for (int cnt = 0; cnt < my_tiles * NU; ++cnt) {
            const int it = cnt / NU, u = cnt - it * NU;
            const int mg = u >> 1, hf = u & 1, buf = it & 1;
            if (u == 0) mbar_wait(go_full(buf), (uint32_t)(it >> 1) & 1u);
            const int stage = cnt % NS;
            mbar_wait(a_full(stage), (uint32_t)(cnt / NS) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t db0 = make_desc_mn_sw128_32b(go0 + (uint32_t)buf * S::GO_BYTES, A_STAGE_BYTES, 512);
                const uint32_t tmem_d = tmem_base + (uint32_t)(mg * C);
                const uint32_t ta = tmem_base + A_COL0 + (uint32_t)(stage * 64);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)      // 8 rows per instruction: two 4-row swizzle atoms = 1024 bytes = +64
                    mma_tf32_ts_mn(tmem_d, ta + (uint32_t)(ks * 8), db0 + (uint64_t)(64 * (hf * 8 + ks)), idesc,
                                   (it == 0 && hf == 0 && ks == 0) ? 0u : 1u);
                mma_commit(a_empty(stage));
                if (u == NU - 1) mma_commit(go_empty(buf));
            }
            __syncwarp();
        }
        if (elect_one()) mma_commit(done_bar);
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// gw += sum over the CTAs of each offset group, gb += sum of the bias slices: a FIXED summation tree (warp w adds slices
// w, w + 8, ... in order, the eight warp sums are added in order), so the result is bit-reproducible.  A warp reads 512
// contiguous bytes of one slice per load and the loads of a thread are independent: a first version that walked all 148
// slices in one thread spent 50 us in a chain of dependent L2 reads.
template <int C>
__global__ void __launch_bounds__(256) k_wgrad_ts_reduce(const float* __restrict__ part, const float* __restrict__ part_b,
                                                         int ctas_per_group, float* __restrict__ gw, float* __restrict__ gb) {
    using S = WtShape<C>;
    constexpr int C4 = C / 4;
    constexpr int total = TS_K * C * C4;
    __shared__ float4 sm[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + lane;
    const float4* src = nullptr;
    float4* dst = nullptr;
    int64_t stride = 0;      // float4 units between consecutive slices
    if (e < total) {
        const int c4 = e % C4, ci = (e / C4) % C, o = e / (C4 * C);
        const int og = o / S::OFFS, ol = o - og * S::OFFS;
        const int mg = ol / S::P, row = (ol % S::P) * S::W + ci;
        src = reinterpret_cast<const float4*>(part + (((int64_t)og * ctas_per_group * S::NMG + mg) * 128 + row) * C) + c4;
        stride = (int64_t)S::NMG * 128 * C4;
        dst = reinterpret_cast<float4*>(gw + ((int64_t)o * C + ci) * C) + c4;
    } else if (e < total + C4 && gb && part_b) {
        src = reinterpret_cast<const float4*>(part_b) + (e - total);
        stride = C4;
        dst = reinterpret_cast<float4*>(gb) + (e - total);
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src) {
#pragma unroll 4
        for (int k = w; k < ctas_per_group; k += 8) {
            const float4 v = __ldg(src + (int64_t)k * stride);
            acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
        }
    }
    sm[w][lane] = acc;
    __syncthreads();
    if (w == 0 && dst) {
        float4 a = sm[0][lane];
#pragma unroll
        for (int j = 1; j < 8; ++j) {
            const float4 v = sm[j][lane];
            a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
        }
        float4 g = *dst;
        g.x += a.x, g.y += a.y, g.z += a.z, g.w += a.w;
        *dst = g;
    }
}

// ------------------------------------------------------------------------------------------------ workspace + launch
struct WgWorkspace {
    float* buf = nullptr;
    size_t bytes = 0;
};
static WgWorkspace& wg_workspace(cudaStream_t st, size_t need, cudaError_t* err) {
    static std::mutex mu;
    static std::unordered_map<cudaStream_t, WgWorkspace> pool;      // one per stream: weight gradients run on side streams
    std::lock_guard<std::mutex> lock(mu);
    WgWorkspace& w = pool[st];
    *err = cudaSuccess;
    if (w.bytes < need) {
        if (w.buf) {
            cudaStreamSynchronize(st);      // earlier launches on this stream may still read it
            cudaFree(w.buf);
        }
        w.buf = nullptr, w.bytes = 0;
        *err = cudaMalloc(&w.buf, need);
        if (*err == cudaSuccess) w.bytes = need;
    }
    return w;
}
static std::atomic<int64_t>* wg_counter() {
    static std::atomic<int64_t> n{0};
    return &n;
}

template <int C>
static int launch_wgrad_ts(WgradTsParams& p, float* gw, float* gb, cudaStream_t stream) {
    using S = WtShape<C>;
    constexpr int MAX_SMEM = 227 * 1024;
    int cap = (MAX_SMEM - S::FIXED_SMEM) / (2 * S::PITCH);
    if (cap > TS_ROWS_CAP) cap = TS_ROWS_CAP;
    cap &= ~7;
    if (cap < 192) return 0;
    p.cap = cap;
    const int smem = S::FIXED_SMEM + 2 * cap * S::PITCH;
    int cpg = sm_count() / S::OG;
    if (cpg > p.book.n_tiles) cpg = p.book.n_tiles;
    p.ctas_per_group = cpg;
    const int grid = cpg * S::OG;
    const size_t part_floats = (size_t)grid * S::NMG * 128 * C, bias_floats = (size_t)cpg * C;
    cudaError_t e;
    WgWorkspace& ws = wg_workspace(stream, (part_floats + bias_floats) * sizeof(float), &e);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("conv_wgrad_ts: workspace of %zu bytes: %s", (part_floats + bias_floats) * sizeof(float), cudaGetErrorString(e));
        return -SCN_ERR_CUDA;
    }
    p.part = ws.buf;
    p.part_b = gb ? ws.buf + part_floats : nullptr;
    auto kern = k_wgrad_ts<C>;
    e = (cudaError_t)ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("conv_wgrad_ts: cudaFuncSetAttribute(%d bytes): %s", smem, cudaGetErrorString(e));
        return -SCN_ERR_CUDA;
    }
    PdlLaunch L(dim3(grid), dim3(S::THREADS), smem, stream);
    e = cudaLaunchKernelEx(&L.cfg, kern, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("conv_wgrad_ts: launch failed: %s", cudaGetErrorString(e));
        return -SCN_ERR_CUDA;
    }
    int rc = check_launch("conv_wgrad_ts");
    if (rc) return -rc;
    k_wgrad_ts_reduce<C><<<(TS_K * C * C / 4 + C / 4 + 31) / 32, 256, 0, stream>>>(p.part, p.part_b, cpg, gw, gb);
    rc = check_launch("conv_wgrad_ts_reduce");
    if (rc) return -rc;
    ++*wg_counter();
    return 1;
}

// 1 = done by the tile-local kernel, 0 = not applicable (caller uses conv_wgrad_tc.cu), < 0 = -status
int conv_wgrad_ts_try(const float* in, int ld_in, int Cin, const int32_t* map, int n_out, int K, const float* go, int ld_go, int Cout,
                      float* gw, float* gb, cudaStream_t stream) {
    if (K != TS_K || !map || Cin != Cout || (Cin != 32 && Cin != 48)) return 0;
    const char* ev = getenv("SCN_WGRAD_TS");      // read per call: tests run both kernels in one process
    if (ev && ev[0] == '0') return 0;
    // C = 48 runs two offset groups with only 2 x 4 gather warps each (TMEM holds 2 A stages beside 7 x 48 accumulator
    // columns) and is instruction-issue bound: 170 us against 128 us for conv_wgrad_tc.cu on the level-0-sized layer
    // (profiles/r2_f_weight_gradient.md).  Parity-tested, but only used on request (SCN_WGRAD_TS=48).
    if (Cin == 48 && !(ev && ev[0] == '4')) return 0;
    if (ld_in % 4 != 0 || ld_go % 4 != 0 || ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(go) | reinterpret_cast<uintptr_t>(gw)) & 15))
        return 0;
    static int min_tiles = -1;
    if (min_tiles < 0) {
        const char* e = getenv("SCN_CONV_TS_MIN_TILES");
        min_tiles = e ? atoi(e) : 2 * sm_count();
    }
    WgradTsParams p;
    if ((n_out + TILE_M - 1) / TILE_M < min_tiles || !tile_book_lookup(map, n_out, &p.book)) return 0;
    p.in = in, p.ld_in = ld_in, p.map = map, p.go = go, p.ld_go = ld_go;
    return Cin == 32 ? launch_wgrad_ts<32>(p, gw, gb, stream) : launch_wgrad_ts<48>(p, gw, gb, stream);
}

}  // namespace scn

extern "C" int64_t scn_conv_wgrad_ts_launch_count(void) { return scn::wg_counter()->load(); }
