// TF32 implicit gather-GEMM on tcgen05 tensor cores (sm_100a).
//
//   out[r, :] = epi( bias + sum_o  in[map[o][r], :] . W[o] )       r in [0, n_out)
//
// One persistent CTA per SM slot owns 128-row output tiles.  The reduction over (kernel offset o,
// 32-channel k-block kb) is a stream of "units"; each unit is one pipeline stage:
//     A stage : 128 gathered input rows x 32 fp32 (16 KB), K-major, 128B-swizzled, written by
//               cp.async (16/8/4-byte, zero-fill for inactive neighbours) from 4 producer warps
//     B stage : W[o][kb] as [Cout_pad x 32] fp32, K-major 128B-swizzled image pre-packed in HBM,
//               fetched with one cp.async.bulk (TMA bulk engine) that completes on the mbarrier
//     MMA     : one thread issues 1..4 tcgen05.mma.kind::tf32 (M=128, N=Cout_pad, K=8), fp32
//               accumulation in TMEM across ALL units of the tile (no read-modify-write of out)
// Accumulators are double buffered in TMEM so the epilogue (tcgen05.ld -> +bias/+residual/ReLU ->
// global) of tile t overlaps the main loop of tile t+1.
//
// Warp roles (9 warps): 0-3 epilogue (TMEM lane quarter = warp id), 4-7 gather producers,
// 8 MMA issuer + TMEM allocator.
#include <cuda.h>   // CUtensorMap types only; the driver entry point is resolved at run time
#include <stdlib.h>
#include <string.h>
#include "tc_common.cuh"

namespace scn {

#ifdef SCN_EXP_TRACE
// timing experiment: globaltimer stamps of CTA 0 at fixed points (slot -> ns), read back with scn_debug_read_trace
__device__ unsigned long long g_trace[32];
__device__ unsigned long long g_cta_times[2 * 512];      // entry / exit stamp of every CTA (first 512)
#define SCN_TRACE(slot)                                                   \
    do {                                                                  \
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {                 \
            unsigned long long t__;                                       \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));       \
            g_trace[slot] = t__;                                          \
        }                                                                 \
    } while (0)
#else
#define SCN_TRACE(slot) \
    do {                \
    } while (0)
#endif

constexpr int CONV_THREADS = 288;
#ifndef SCN_CONV_TAILSPLIT_DEFAULT
#define SCN_CONV_TAILSPLIT_DEFAULT 1
#endif

struct ConvTcParams {
    const float* in;
    int ld_in, Cin;
    const int32_t* map;
    int n_out, K;
    const uint8_t* image;
    const float* bias;
    const float* residual;
    int ld_res;
    const float* mask;
    int ld_mask;
    float* out;
    int ld_out, Cout, epi;
    float* out2;          // optional second output: out2 = epi2(out) (RELU and / or ROUND), the next layer's operand
    int ld_out2, epi2;
    int cout_pad, n_kb, cin_pad8, stages, tmem_cols, n_tiles;
    int osplit, opg;      // offsets of a tile are split over `osplit` work items of `opg` offsets each (small levels)
    float* scratch;       // split mode: accumulation buffer [n_out][Cout], all-zero between launches
    unsigned int* tickets;  // split mode: one self-resetting arrival counter per tile
    int cluster;          // > 1: the osplit work items of a tile are one thread-block cluster and reduce through DSMEM
    int skip;             // 1: producers skip rows that are inactive now and were inactive in the stage's previous use
    int n_whole, n_work;  // work items [0, n_whole) are whole tiles (all K offsets, direct epilogue); items beyond are the
                          // offset groups of tiles n_whole, n_whole + 1, ... (osplit per tile).  n_whole = 0: every tile is
                          // split; n_whole = n_tiles: none is; in between: only the tail tiles of the last wave are
};

// work item -> (tile, offset range)
__device__ __forceinline__ void decode_item(const ConvTcParams& p, int w, int& tile, int& o_lo, int& o_hi) {
    if (w < p.n_whole) {
        tile = w, o_lo = 0, o_hi = p.K;
    } else {
        const int v = w - p.n_whole, t = v / p.osplit;
        tile = p.n_whole + t, o_lo = (v - t * p.osplit) * p.opg, o_hi = min(p.K, o_lo + p.opg);
    }
}

// TMA = true : the A stage is filled by ONE producer warp with `cp.async.bulk.tensor.2d ... tile::gather4`
//              (SASS UTMALDG): lane L gathers rows 4L..4L+3 of the tile by index, the hardware applies the
//              128B swizzle, zero-fills index -1 / out-of-range columns, rounds fp32 -> TF32 (to nearest)
//              and completes the bytes on the stage mbarrier.  6 warps.
// TMA = false: cp.async producers (4 warps, 16/8/4-byte copies) for feature tensors whose row stride is not
//              a multiple of 16 bytes.  9 warps.
template <int VEC, bool TMA>
__global__ void __launch_bounds__(TMA ? 192 : CONV_THREADS, TMA ? 2 : 3) k_conv_tc(const __grid_constant__ CUtensorMap tmap,
                                                                        const ConvTcParams p) {
    constexpr int MMA_WARP = TMA ? 5 : 8;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int S = p.stages;
    const uint32_t stage_bytes = A_STAGE_BYTES + (uint32_t)p.cout_pad * 128u;
    const uint32_t bars = smem_base + (uint32_t)S * stage_bytes;  // 8-byte barriers
    // layout: full[S], empty[S], acc_full[2], acc_empty[2], tmem_ptr
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
    auto accf_bar = [&](int b) { return bars + 8u * (2 * S + b); };
    auto acce_bar = [&](int b) { return bars + 8u * (2 * S + 2 + b); };
    const uint32_t tmem_slot = bars + 8u * (2 * S + 4);

    // warp index through a shuffle from lane 0: the compiler then knows it is warp-uniform (uniform registers for the ring
    // bookkeeping and barrier addresses, no vote loops around barrier instructions; cutlass::canonical_warp_idx_sync)
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int units_per_tile = p.K * p.n_kb;
    if (tid == 0) SCN_TRACE(0);
#ifdef SCN_EXP_TRACE
    if (tid == 0 && blockIdx.x < 512) {
        unsigned long long t__;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));
        g_cta_times[2 * blockIdx.x] = t__;
    }
#endif

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), TMA ? 1 : 32 + 1);      // cp.async path: the 32 lanes of the owning warp + expect_tx
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(accf_bar(b), 1);
            mbar_init(acce_bar(b), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (tid == 0) SCN_TRACE(1);
    // PDL: everything above touched only shared memory / TMEM / parameters; the next kernel on the stream may start its
    // own prologue now, and we wait here for the previous kernel's results
    pdl_trigger();
    pdl_wait();

    if (TMA && warp == 4) {
        // ===================== TMA gather producer (one warp) =====================
        if constexpr (TMA) {
            int s = 0;
            uint32_t ph = 0;
            auto load_idx = [&](int tile, int o, int (&dst)[4]) {
                const int row0 = tile * TILE_M + 4 * lane;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    int r = row0 + i;
                    int v = -1;
                    if (tile < p.n_tiles && r < p.n_out) v = p.map ? __ldg(p.map + (int64_t)o * p.n_out + r) : r;
                    dst[i] = v;
                }
            };
            int idx[4], idx_next[4];
            const int n_work = p.n_work;
            {
                int t0_, lo0_, hi0_;
                decode_item(p, blockIdx.x, t0_, lo0_, hi0_);
                load_idx(t0_, lo0_, idx_next);
            }
            const uint32_t wbytes = (uint32_t)p.cout_pad * 128u;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                int tile, o_lo, o_hi;
                decode_item(p, w, tile, o_lo, o_hi);
                for (int o = o_lo; o < o_hi; ++o) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) idx[i] = idx_next[i];
                    if (o + 1 < o_hi) load_idx(tile, o + 1, idx_next);
                    else {
                        const int wn = w + gridDim.x;
                        int tn = p.n_tiles, lon = 0, hin = 0;
                        if (wn < n_work) decode_item(p, wn, tn, lon, hin);
                        load_idx(tn, lon, idx_next);
                    }
                    for (int kb = 0; kb < p.n_kb; ++kb) {
                        mbar_wait(empty_bar(s), ph ^ 1);
                        const uint32_t a_stage = smem_base + (uint32_t)s * stage_bytes;
                        if (lane == 0) {
                            mbar_arrive_expect_tx(full_bar(s), (uint32_t)A_STAGE_BYTES + wbytes);
                            bulk_g2s(a_stage + A_STAGE_BYTES, p.image + (size_t)(o * p.n_kb + kb) * wbytes, wbytes,
                                     full_bar(s));
                        }
                        asm volatile(
                            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
                            "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(a_stage + (uint32_t)lane * 512u),
                            "l"(&tmap), "r"(kb * KB), "r"(idx[0]), "r"(idx[1]), "r"(idx[2]), "r"(idx[3]), "r"(full_bar(s))
                            : "memory");
                        if (++s == S) s = 0, ph ^= 1;
                    }
                }
            }
        }
    } else if (!TMA && warp >= 4 && warp < 8) {
        // ===================== gather producers =====================
        // Each producer warp OWNS every fourth unit of the CTA's flat unit stream (unit q = work item, offset, k-block in
        // consumption order; warp pw fills q = pw, pw + 4, ...): all 128 rows of the stage, its weight block, and the
        // only arrivals on its full barrier.  Why: the producers are bound by the LENGTH of their own serial
        // instruction stream, not by memory (profiles/r1_h_issue_bound.md: with copies, MMAs, weight fetches and map
        // reads all compiled out a unit still cost ~800 cycles per CTA).  When all four warps cooperate on every unit,
        // each pays the per-unit overhead (barrier probe, ring bookkeeping, loop) for every unit; with owned units that
        // chain is paid once per four units and four units are in production concurrently.
        // Lane L reads the neighbour indices of rows L, L+32, L+64, L+96 (coalesced), one own unit (= four units)
        // ahead; the copy of row (lane >> 3) + 4 i fetches its index with a shuffle, 8 lanes x 16 bytes per row.
        // A warp that waits for "round k-1 of stage s consumed" can only tell it from parity, so round k-2 must already be
        // known consumed when it looks: with G owner warps that holds iff S >= G (the warp's previous unit q-G needed unit
        // q-G-S consumed, and q-2S <= q-G-S).  Hence G = min(4, S) owners; with S = 3 the fourth producer warp idles.
        const int G = S < 4 ? S : 4;
        const int pw = warp - 4;
        const int c = lane & 7, rsub = lane >> 3;
        const uint32_t dst_even = (uint32_t)rsub * 128u + (uint32_t)((c ^ rsub) << 4);
        const uint32_t dst_odd = (uint32_t)rsub * 128u + (uint32_t)((c ^ (rsub + 4)) << 4);
        const int n_work = p.n_work;
        const int wstep = gridDim.x;
        const uint32_t wbytes = (uint32_t)p.cout_pad * 128u;
        (void)wbytes;
        const uint32_t row_bytes = (uint32_t)p.ld_in * 4u;
        const char* in_c = reinterpret_cast<const char*>(p.in) + c * 16;

        struct Cursor {
            int w, o, o_hi, kb, row0;      // row0 = first row of the tile
        };
        auto enter = [&](Cursor& q) {      // a division per WORK ITEM, not per unit
            int tile;
            decode_item(p, q.w, tile, q.o, q.o_hi);
            q.row0 = tile * TILE_M;
            q.kb = 0;
        };
        auto step = [&](Cursor& q) {
            if (++q.kb == p.n_kb) {
                q.kb = 0;
                if (++q.o >= q.o_hi) {
                    q.w += wstep;
                    enter(q);
                }
            }
        };
        auto load_idx = [&](const Cursor& q, int (&dst)[4]) {
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = -1;
            if (q.w < n_work) {
                const int r0 = q.row0 + lane;
#ifdef SCN_EXP_NOIDX
                const int32_t* mp = nullptr;
#else
                const int32_t* mp = p.map ? p.map + (int64_t)q.o * p.n_out + r0 : nullptr;
#endif
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (r0 + 32 * j < p.n_out) dst[j] = mp ? __ldg(mp + 32 * j) : r0 + 32 * j;
            }
        };
        Cursor cur, nxt;
        cur.w = blockIdx.x;
        enter(cur);
        for (int j = 0; j < pw; ++j) step(cur);
        nxt = cur;
        for (int j = 0; j < G; ++j) step(nxt);
        int s = pw;                                   // pw < G <= S
        uint32_t ph = 0;
        int idx[4], idx_next[4];
        // Row skipping (profiles/r1_j: the shared-memory fill path is the unit closest to its ceiling and 63 % of the rows
        // it writes are zero fill): with S == G an owner warp always refills the SAME stage, so lane L knows from its
        // previous unit whether rows L + 32 j of that stage already hold zeros.  Bit j of `dirty` = "row may be
        // non-zero"; an inactive row over a clean row needs no copy at all (code -2: the cp.async is predicated off),
        // an inactive row over a dirty row is zero-filled (code -1, ignore-src), an active row is copied.
        const bool skipping = p.skip && G == S;
        uint32_t dirty = 0xFu;                       // shared memory starts undefined
        load_idx(cur, idx);
        if (pw == 0 && idx[0] == -12345) SCN_TRACE(31);
        if (pw == 0) SCN_TRACE(2);
        while (pw < G && cur.w < n_work) {
            load_idx(nxt, idx_next);
            mbar_wait(empty_bar(s), ph ^ 1);
            const uint32_t a_stage = smem_base + (uint32_t)s * stage_bytes;
            const uint32_t fb = full_bar(s);
            if (elect_one()) {
#ifdef SCN_EXP_NOWEIGHT
                mbar_arrive(fb);
#else
                mbar_arrive_expect_tx(fb, wbytes);
                bulk_g2s(a_stage + A_STAGE_BYTES, p.image + (size_t)(cur.o * p.n_kb + cur.kb) * wbytes, wbytes, fb);
#endif
            }
            const int col0 = cur.kb * KB + c * 4;
            // shuffles need the whole warp: lane-dependent conditions only predicate the copies
            if constexpr (VEC == 4) {
                // one ISETP + IMAD.WIDE + LDGSTS per 16-byte chunk; inactive rows use the ignore-src form
                // (zero fill, the address is never dereferenced, so index -1 needs no clamp)
                const char* colp = in_c + cur.kb * (KB * 4);
                const uint32_t de = a_stage + dst_even, dodd = a_stage + dst_odd;
                if (p.Cin - cur.kb * KB >= KB) {      // warp-uniform: a full 32-channel block
                    uint32_t now = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int code = idx[j] >= 0 ? idx[j] : (((dirty >> j) & 1u) ? -1 : -2);
                        now |= (idx[j] >= 0 ? 1u : 0u) << j;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
#ifdef SCN_EXP_NOCOPY
                            if (p.n_out > 0) break;
#endif
                            const int r = __shfl_sync(0xffffffffu, code, rsub + 4 * i);
                            const char* src = colp + (uint64_t)(uint32_t)r * row_bytes;
                            asm volatile(
                                "{\n\t"
                                ".reg .pred p, q;\n\t"
                                "setp.lt.s32 p, %2, 0;\n\t"
                                "setp.ne.s32 q, %2, -2;\n\t"
                                "@q cp.async.cg.shared.global [%0], [%1], 16, p;\n\t"
                                "}" ::"r"(((i & 1) ? dodd : de) + (uint32_t)(32 * j + 4 * i) * 128u),
                                "l"(src), "r"(r)
                                : "memory");
                        }
                    }
                    if (skipping) dirty = now;
                } else {
                    // last, partial block: the MMA reads chunks below cin_pad8 ("wanted"), chunks at or beyond Cin must be
                    // zeros, lanes of the other chunks do nothing.  Row skipping must not lengthen this loop (an earlier
                    // version carried the dirty bit of active rows through the shuffle and let the idle lanes clear their
                    // chunks: C = 16 layers ran 1.8x slower, profiles/r1_p): same three-way code as a full block; the price
                    // is that a partial unit cannot make a row CLEAN for a later full block in the same stage (chunks the
                    // MMA does not read keep stale bytes), so it only ever sets dirty bits -- unless every unit of the
                    // layer is this block (n_kb == 1), where those chunks are never read at all.
                    // One lane-constant clamp instead of branches around the copy (an `if (wanted)` around inline PTX that
                    // carries its own guard predicate compiles to a divergent branch + BSSY/BSYNC per copy): a real chunk
                    // keeps its code, a zero-padding chunk (Cin <= column < cin_pad8) is zero-filled unless the row is
                    // clean (-1), a chunk the MMA does not read is always skipped (-2).
                    uint32_t now = 0;
                    if (p.cin_pad8 - cur.kb * KB <= 16) {
                        // at most four chunks of a row are read (C = 16, 48, 80, 112: the last 16 channels): four lanes per
                        // row, EIGHT rows per copy instruction -- 16 copies per unit instead of 32 with half the lanes idle
                        const int c4 = lane & 3, r8 = lane >> 2;
                        const int col4 = cur.kb * KB + c4 * 4;
                        const int clamp4 = col4 < p.cin_pad8 ? (col4 < p.Cin ? 0x7fffffff : -1) : -2;
                        const char* colp4 = reinterpret_cast<const char*>(p.in) + cur.kb * (KB * 4) + c4 * 16;
                        const uint32_t d8 = a_stage + (uint32_t)r8 * 128u + (uint32_t)((c4 ^ r8) << 4);      // row & 7 == r8
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int code = idx[j] >= 0 ? idx[j] : (((dirty >> j) & 1u) ? -1 : -2);
                            now |= (idx[j] >= 0 ? 1u : 0u) << j;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int r = min(__shfl_sync(0xffffffffu, code, r8 + 8 * i), clamp4);
                                const char* src = colp4 + (uint64_t)(uint32_t)r * row_bytes;
                                asm volatile(
                                    "{\n\t"
                                    ".reg .pred p, q;\n\t"
                                    "setp.lt.s32 p, %2, 0;\n\t"
                                    "setp.ne.s32 q, %2, -2;\n\t"
                                    "@q cp.async.cg.shared.global [%0], [%1], 16, p;\n\t"
                                    "}" ::"r"(d8 + (uint32_t)(32 * j + 8 * i) * 128u),
                                    "l"(src), "r"(r)
                                    : "memory");
                            }
                        }
                    } else {
                    const int clamp = col0 < p.cin_pad8 ? (col0 < p.Cin ? 0x7fffffff : -1) : -2;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int code = idx[j] >= 0 ? idx[j] : (((dirty >> j) & 1u) ? -1 : -2);
                        now |= (idx[j] >= 0 ? 1u : 0u) << j;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = min(__shfl_sync(0xffffffffu, code, rsub + 4 * i), clamp);
                            const char* src = colp + (uint64_t)(uint32_t)r * row_bytes;
                            asm volatile(
                                "{\n\t"
                                ".reg .pred p, q;\n\t"
                                "setp.lt.s32 p, %2, 0;\n\t"
                                "setp.ne.s32 q, %2, -2;\n\t"
                                "@q cp.async.cg.shared.global [%0], [%1], 16, p;\n\t"
                                "}" ::"r"(((i & 1) ? dodd : de) + (uint32_t)(32 * j + 4 * i) * 128u),
                                "l"(src), "r"(r)
                                : "memory");
                        }
                    }
                    }
                    if (skipping) dirty = p.n_kb == 1 ? now : (dirty | now);
                }
            } else {
                // 8- and 4-byte copies (row stride not a multiple of 16 bytes): each 16-byte chunk is 4 / VEC sub-copies,
                // every one the same guarded copy as above behind its own lane-constant clamp (real column: the row's
                // code; zero padding inside a chunk the MMA reads: zero fill unless the row is clean; chunk the MMA does not
                // read: skip).  Branch-free: the mask network's 22-channel layers run here.
                constexpr int NSUB = 4 / VEC;
                int clamp[NSUB];
#pragma unroll
                for (int u = 0; u < NSUB; ++u)
                    clamp[u] = col0 + u * VEC < p.Cin ? 0x7fffffff : (col0 < p.cin_pad8 ? -1 : -2);
                const bool clears = p.n_kb == 1 || p.cin_pad8 - cur.kb * KB >= KB;      // a full block rewrites every chunk
                const char* colp = in_c + cur.kb * (KB * 4);
                const uint32_t de = a_stage + dst_even, dodd = a_stage + dst_odd;
                uint32_t now = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int code = idx[j] >= 0 ? idx[j] : (((dirty >> j) & 1u) ? -1 : -2);
                    now |= (idx[j] >= 0 ? 1u : 0u) << j;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r0 = __shfl_sync(0xffffffffu, code, rsub + 4 * i);
                        const char* src = colp + (uint64_t)(uint32_t)r0 * row_bytes;
                        const uint32_t dst = ((i & 1) ? dodd : de) + (uint32_t)(32 * j + 4 * i) * 128u;
#pragma unroll
                        for (int u = 0; u < NSUB; ++u) {
                            const int r = min(r0, clamp[u]);
                            if constexpr (VEC == 2)
                                asm volatile(
                                    "{\n\t"
                                    ".reg .pred p, q;\n\t"
                                    "setp.lt.s32 p, %2, 0;\n\t"
                                    "setp.ne.s32 q, %2, -2;\n\t"
                                    "@q cp.async.ca.shared.global [%0], [%1], 8, p;\n\t"
                                    "}" ::"r"(dst + (uint32_t)u * 8u),
                                    "l"(src + u * 8), "r"(r)
                                    : "memory");
                            else
                                asm volatile(
                                    "{\n\t"
                                    ".reg .pred p, q;\n\t"
                                    "setp.lt.s32 p, %2, 0;\n\t"
                                    "setp.ne.s32 q, %2, -2;\n\t"
                                    "@q cp.async.ca.shared.global [%0], [%1], 4, p;\n\t"
                                    "}" ::"r"(dst + (uint32_t)u * 4u),
                                    "l"(src + u * 4), "r"(r)
                                    : "memory");
                        }
                    }
                }
                if (skipping) dirty = clears ? now : (dirty | now);
            }
            // the stage's full barrier receives this thread's arrival when its copies have landed
            cp_async_mbar_arrive_noinc(fb);
            if (pw == 0) SCN_TRACE(3);
            s += G;
            if (s >= S) s -= S, ph ^= 1;
            cur = nxt;
#pragma unroll
            for (int j = 0; j < 4; ++j) idx[j] = idx_next[j];
            for (int j = 0; j < G; ++j) step(nxt);
        }
        cp_async_wait_all();
    } else if (warp == MMA_WARP) {
        // ===================== MMA issuer =====================
        // The whole warp stays converged and waits; one elected lane issues (cute::elect_one_sync pattern).  From a
        // divergent `if (lane == 0)` region every UTCHMMA / UTCBAR is wrapped in an ELECT + BRA.U.ANY vote loop and the
        // descriptors are rebuilt with ~20 uniform-datapath instructions per unit -- with owned producer units the
        // issuing thread's serial instruction stream is what bounds the CTA (profiles/r1_h_issue_bound.md).
        const uint32_t idesc = make_idesc_tf32(TILE_M, p.cout_pad);
        const uint64_t desc0 = make_desc_sw128(smem_base);      // stage 0, A operand; the address field counts 16-byte units
        const uint32_t stage_d = stage_bytes >> 4;
        int s = 0, it = 0;
        uint32_t ph = 0;
        const int n_work = p.n_work;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
            int tile_, o_lo, o_hi;
            decode_item(p, w, tile_, o_lo, o_hi);
            const int b = it & 1;
            mbar_wait(acce_bar(b), ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(b * p.cout_pad);
            uint32_t accum = 0;
            for (int o = o_lo; o < o_hi; ++o) {
                for (int kb = 0; kb < p.n_kb; ++kb) {
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    if (accum == 0) SCN_TRACE(4);
                    if (elect_one()) {
                        const uint64_t da = desc0 + (uint64_t)((uint32_t)s * stage_d);
                        const uint64_t db = da + (uint64_t)(A_STAGE_BYTES >> 4);
                        const int kcols = min(KB, p.cin_pad8 - kb * KB);
#ifndef SCN_EXP_NOMMA
                        // advance 32 bytes (8 tf32) inside the 128-byte swizzled row: +2 in the >>4 address field
                        mma_tf32(tmem_d, da, db, idesc, accum);
                        if (kcols > 8) mma_tf32(tmem_d, da + 2, db + 2, idesc, 1u);
                        if (kcols > 16) mma_tf32(tmem_d, da + 4, db + 4, idesc, 1u);
                        if (kcols > 24) mma_tf32(tmem_d, da + 6, db + 6, idesc, 1u);
#endif
                        mma_commit(empty_bar(s));
                    }
                    accum = 1u;
                    if (++s == S) s = 0, ph ^= 1;
                }
            }
            if (elect_one()) mma_commit(accf_bar(b));
            SCN_TRACE(5);
        }
        (void)units_per_tile;
    } else if (warp < 4) {
        // ===================== epilogue warps 0..3 =====================
        int it = 0;
        const int epi = p.epi;
        const bool vec_ok = (p.ld_out % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                            (!p.out2 || ((p.ld_out2 % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out2) & 15) == 0))) &&
                            (!(epi & SCN_EPI_ADD) ||
                             ((p.ld_res % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0))) &&
                            (!(epi & SCN_EPI_MASK) ||
                             ((p.ld_mask % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.mask) & 15) == 0)));
        const uint32_t part_pitch = (uint32_t)p.cout_pad * 4u + 16u;      // cluster mode: row pitch of the partial tile
        // order: bias, MASK (aux > 0 ? x : 0), ADD residual, RELU, ROUND (rna to TF32)
        auto finish = [&](float x, float m, float r) {
            if ((epi & SCN_EPI_MASK) && !(m > 0.f)) x = 0.f;
            if (epi & SCN_EPI_ADD) x += r;
            if (epi & SCN_EPI_RELU) x = fmaxf(x, 0.f);
            if (epi & SCN_EPI_ROUND) {
                uint32_t t;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x));
                x = __uint_as_float(t);
            }
            return x;
        };
        const int n_work = p.n_work;
        __shared__ int s_last;
        // bias / mask / residual / ReLU / rounding of 16 accumulator columns of one row, then the store
        auto finish_store = [&](float (&v)[16], int row, int c0) {
            float* orow = p.out + (int64_t)row * p.ld_out + c0;
            float* orow2 = p.out2 ? p.out2 + (int64_t)row * p.ld_out2 + c0 : nullptr;
            const float* rrow = (epi & SCN_EPI_ADD) ? p.residual + (int64_t)row * p.ld_res + c0 : nullptr;
            const float* mrow = (epi & SCN_EPI_MASK) ? p.mask + (int64_t)row * p.ld_mask + c0 : nullptr;
            if (p.bias) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < p.Cout) v[j] += __ldg(p.bias + c0 + j);
            }
            if (vec_ok && c0 + 16 <= p.Cout) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 r4 = rrow ? *reinterpret_cast<const float4*>(rrow + j) : make_float4(0, 0, 0, 0);
                    float4 m4 = mrow ? *reinterpret_cast<const float4*>(mrow + j) : make_float4(1, 1, 1, 1);
                    float4 x;
                    x.x = finish(v[j], m4.x, r4.x), x.y = finish(v[j + 1], m4.y, r4.y);
                    x.z = finish(v[j + 2], m4.z, r4.z), x.w = finish(v[j + 3], m4.w, r4.w);
                    *reinterpret_cast<float4*>(orow + j) = x;
                    if (orow2) *reinterpret_cast<float4*>(orow2 + j) = epi2_apply4(x, p.epi2);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < p.Cout) {
                        const float y = finish(v[j], mrow ? mrow[j] : 1.f, rrow ? rrow[j] : 0.f);
                        orow[j] = y;
                        if (orow2) orow2[j] = epi2_apply(y, p.epi2);
                    }
            }
        };
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
            int tile, o_lo_, o_hi_;
            decode_item(p, w, tile, o_lo_, o_hi_);
            const int b = it & 1;
            mbar_wait<200>(accf_bar(b), (it >> 1) & 1);      // epilogue warps wait a whole tile: long back-off
            tc_fence_after();
            if (warp == 0) SCN_TRACE(6);
            const int row = tile * TILE_M + warp * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(b * p.cout_pad);
            if (w < p.n_whole) {
                for (int c0 = 0; c0 < p.cout_pad; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + c0, v);
                    if (row < p.n_out && c0 < p.Cout) finish_store(v, row, c0);
                }
                tc_fence_before();
                mbar_arrive(acce_bar(b));
            } else if (p.cluster > 1) {
                // cluster mode: the partial sum of this offset group goes to this CTA's own shared memory (the stage ring is
                // free: every MMA of the work item has completed); the cluster reduces it through DSMEM below
                const uint32_t prow = smem_base + (uint32_t)(warp * 32 + lane) * part_pitch;
                for (int c0 = 0; c0 < p.cout_pad; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + c0, v);
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(prow + (uint32_t)(c0 + j) * 4u), "f"(v[j]),
                                     "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3])
                                     : "memory");
                }
                tc_fence_before();
                mbar_arrive(acce_bar(b));
            } else {
                // split-offset mode: this work item holds the PARTIAL sum of its offset group.  Partials are added with
                // fp32 atomics into a library-owned accumulation buffer that is all-zero between launches; the last group
                // of the tile to arrive (ticket counter) reads the complete sums, applies the epilogue, writes the output
                // and zeroes the buffer again: one launch, no prefill / post-pass kernels.
                float* arow = p.scratch + (int64_t)row * p.Cout;
                const bool c4 = (p.Cout & 3) == 0;
                for (int c0 = 0; c0 < p.cout_pad; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + c0, v);      // warp-collective: every lane, also rows beyond n_out
                    if (row < p.n_out) {
                        if (c4) {      // 16-byte vector reductions: a quarter of the atomic operations
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                if (c0 + j < p.Cout)
                                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(arow + c0 + j),
                                                 "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3])
                                                 : "memory");
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (c0 + j < p.Cout) atomicAdd(arow + c0 + j, v[j]);
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(acce_bar(b));
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (tid == 0) {
                    const unsigned int t = atomicAdd(p.tickets + tile, 1u);
                    s_last = (t == (unsigned int)p.osplit - 1u);
                    if (s_last) p.tickets[tile] = 0;      // ready for the next launch on this stream
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (warp == 0) SCN_TRACE(9);
                if (s_last) {
                    // final pass over the tile as a flat [rows x Cout] array: consecutive threads take consecutive
                    // elements (coalesced; thread-per-row scalar accesses made this pass 20x slower)
                    __threadfence();
                    const int row0 = tile * TILE_M;
                    const int rows = min(TILE_M, p.n_out - row0);
                    float* abase = p.scratch + (int64_t)row0 * p.Cout;
                    auto fin1 = [&](float x, int r, int c) {
                        if (p.bias) x += __ldg(p.bias + c);
                        const float m = (epi & SCN_EPI_MASK) ? p.mask[(int64_t)r * p.ld_mask + c] : 1.f;
                        const float rs = (epi & SCN_EPI_ADD) ? p.residual[(int64_t)r * p.ld_res + c] : 0.f;
                        return finish(x, m, rs);
                    };
                    if (c4 && vec_ok) {
                        // four 16-byte loads in flight per thread: the pass is a chain of L2 round trips otherwise
                        const int q = p.Cout >> 2, total = rows * q;
                        float4* ap = reinterpret_cast<float4*>(abase);
                        for (int e0 = tid; e0 < total; e0 += 4 * 128) {
                            float4 x[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (e0 + 128 * u < total) x[u] = __ldcg(ap + e0 + 128 * u);
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (e0 + 128 * u < total) __stcg(ap + e0 + 128 * u, make_float4(0.f, 0.f, 0.f, 0.f));
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int e = e0 + 128 * u;
                                if (e < total) {
                                    const int r = e / q, c = (e - r * q) << 2, gr = row0 + r;
                                    float4 y;
                                    y.x = fin1(x[u].x, gr, c), y.y = fin1(x[u].y, gr, c + 1);
                                    y.z = fin1(x[u].z, gr, c + 2), y.w = fin1(x[u].w, gr, c + 3);
                                    *reinterpret_cast<float4*>(p.out + (int64_t)gr * p.ld_out + c) = y;
                                    if (p.out2) *reinterpret_cast<float4*>(p.out2 + (int64_t)gr * p.ld_out2 + c) = epi2_apply4(y, p.epi2);
                                }
                            }
                        }
                    } else {
                        const int total = rows * p.Cout;
                        for (int e = tid; e < total; e += 128) {
                            const int r = e / p.Cout, c = e - r * p.Cout;
                            const float x = __ldcg(abase + e);
                            __stcg(abase + e, 0.f);
                            const float y = fin1(x, row0 + r, c);
                            p.out[(int64_t)(row0 + r) * p.ld_out + c] = y;
                            if (p.out2) p.out2[(int64_t)(row0 + r) * p.ld_out2 + c] = epi2_apply(y, p.epi2);
                        }
                    }
                }
                if (warp == 0) SCN_TRACE(10);
                asm volatile("bar.sync 1, 128;" ::: "memory");      // s_last is rewritten by the next work item
            }
        }
    }

    if (p.cluster > 1) {
        // ===================== cluster reduction through distributed shared memory =====================
        // Work item w = tile * osplit + g is CTA g of cluster `tile`.  After the cluster barrier every CTA adds, for its
        // slice of the tile's rows, the osplit partial tiles in rank order (deterministic, no global atomics, no scratch),
        // applies the epilogue and writes the output rows.
        cluster_sync_all();
        if (warp < 4) {
            const int cs = p.cluster;
            const int tile = blockIdx.x / cs;
            const int rank = (int)cluster_ctarank();
            const int row0 = tile * TILE_M;
            const int rows = min(TILE_M, p.n_out - row0);
            const int per = (TILE_M + cs - 1) / cs;
            const int r_lo = rank * per, r_hi = min(rows, r_lo + per);
            const uint32_t part_pitch = (uint32_t)p.cout_pad * 4u + 16u;
            const int epi = p.epi;
            uint32_t rbase[8];
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) rbase[rr] = mapa_shared(smem_base, (uint32_t)(rr < cs ? rr : 0));
            auto fin1 = [&](float x, int r, int c) {
                if (p.bias) x += __ldg(p.bias + c);
                if ((epi & SCN_EPI_MASK) && !(p.mask[(int64_t)r * p.ld_mask + c] > 0.f)) x = 0.f;
                if (epi & SCN_EPI_ADD) x += p.residual[(int64_t)r * p.ld_res + c];
                if (epi & SCN_EPI_RELU) x = fmaxf(x, 0.f);
                if (epi & SCN_EPI_ROUND) {
                    uint32_t t;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x));
                    x = __uint_as_float(t);
                }
                return x;
            };
            if (r_hi > r_lo) {
                const bool v4 = (p.Cout & 3) == 0 && (p.ld_out & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 &&
                                (!p.out2 || ((p.ld_out2 & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out2) & 15) == 0));
                if (v4) {
                    const int q = p.Cout >> 2, total = (r_hi - r_lo) * q;
                    for (int e = tid; e < total; e += 128) {
                        const int r = r_lo + e / q, c = (e % q) << 2;
                        const uint32_t off = (uint32_t)r * part_pitch + (uint32_t)c * 4u;
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr)
                            if (rr < cs) {
                                const float4 v = ld_dsmem_v4(rbase[rr] + off);
                                acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
                            }
                        const int gr = row0 + r;
                        acc.x = fin1(acc.x, gr, c), acc.y = fin1(acc.y, gr, c + 1), acc.z = fin1(acc.z, gr, c + 2);
                        acc.w = fin1(acc.w, gr, c + 3);
                        *reinterpret_cast<float4*>(p.out + (int64_t)gr * p.ld_out + c) = acc;
                        if (p.out2) *reinterpret_cast<float4*>(p.out2 + (int64_t)gr * p.ld_out2 + c) = epi2_apply4(acc, p.epi2);
                    }
                } else {
                    const int total = (r_hi - r_lo) * p.Cout;
                    for (int e = tid; e < total; e += 128) {
                        const int r = r_lo + e / p.Cout, c = e % p.Cout;
                        const uint32_t off = (uint32_t)r * part_pitch + (uint32_t)c * 4u;
                        float acc = 0.f;
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr)
                            if (rr < cs) acc += ld_dsmem_f32(rbase[rr] + off);
                        acc = fin1(acc, row0 + r, c);
                        p.out[(int64_t)(row0 + r) * p.ld_out + c] = acc;
                        if (p.out2) p.out2[(int64_t)(row0 + r) * p.ld_out2 + c] = epi2_apply(acc, p.epi2);
                    }
                }
            }
        }
        cluster_sync_all();      // nobody leaves while its shared memory is still being read
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) SCN_TRACE(11);
#ifdef SCN_EXP_TRACE
    if (tid == 0 && blockIdx.x < 512) {
        unsigned long long t__;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));
        g_cta_times[2 * blockIdx.x + 1] = t__;
    }
#endif
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// pack W_eff[o][ci][co] into the kernel's B-stage images: [K][n_kb][cout_pad][32] fp32, element
// (n, k) of a block at byte  n*128 + (((k>>2) ^ (n&7)) << 4) + (k&3)*4  (K-major SWIZZLE_128B).
__global__ void k_pack_weights(const float* __restrict__ w, int K, int A, int B, int transpose, int reverse, int Cin,
                               int Cout, int cout_pad, int n_kb, float* __restrict__ image) {
    int64_t total = (int64_t)K * n_kb * cout_pad * KB;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int k = (int)(i % KB);
        int n = (int)((i / KB) % cout_pad);
        int kb = (int)((i / ((int64_t)KB * cout_pad)) % n_kb);
        int o = (int)(i / ((int64_t)KB * cout_pad * n_kb));
        int ci = kb * KB + k;
        float v = 0.f;
        if (ci < Cin && n < Cout) {
            int oo = reverse ? K - 1 - o : o;
            const float* pw = w + (int64_t)oo * A * B;
            v = transpose ? pw[(int64_t)n * B + ci] : pw[(int64_t)ci * B + n];
            uint32_t t;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
            v = __uint_as_float(t);
        }
        int64_t blk = ((int64_t)o * n_kb + kb) * cout_pad * KB;
        int off = n * KB + ((((k >> 2) ^ (n & 7)) << 2) | (k & 3));
        image[blk + off] = v;
    }
}

// one launch for many images: blockIdx.y = table row, blockIdx.x strides over the image
__global__ void k_pack_weights_multi(const int64_t* __restrict__ table) {
    const int64_t* e = table + (int64_t)blockIdx.y * 7;
    const float* w = reinterpret_cast<const float*>(e[0]);
    float* image = reinterpret_cast<float*>(e[1]);
    const int K = (int)e[2], Cin = (int)e[3], Cout = (int)e[4], transpose = (int)e[5], reverse = (int)e[6];
    const int A = transpose ? Cout : Cin, B = transpose ? Cin : Cout;
    const int cout_pad = (Cout + 15) / 16 * 16, n_kb = (Cin + KB - 1) / KB;
    const int64_t total = (int64_t)K * n_kb * cout_pad * KB;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int k = (int)(i % KB);
        int n = (int)((i / KB) % cout_pad);
        int kb = (int)((i / ((int64_t)KB * cout_pad)) % n_kb);
        int o = (int)(i / ((int64_t)KB * cout_pad * n_kb));
        int ci = kb * KB + k;
        float v = 0.f;
        if (ci < Cin && n < Cout) {
            int oo = reverse ? K - 1 - o : o;
            const float* pw = w + (int64_t)oo * A * B;
            v = transpose ? pw[(int64_t)n * B + ci] : pw[(int64_t)ci * B + n];
            uint32_t t;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
            v = __uint_as_float(t);
        }
        int64_t blk = ((int64_t)o * n_kb + kb) * cout_pad * KB;
        int off = n * KB + ((((k >> 2) ^ (n & 7)) << 2) | (k & 3));
        image[blk + off] = v;
    }
}

}  // namespace scn

using namespace scn;

// Split-offset mode workspace: accumulation buffer (kept all-zero between launches by the kernel itself) + per-tile
// ticket counters, owned by the library, grow-only (a reallocation synchronises the device once; the sizes settle after
// the first step).  One workspace PER STREAM (kernels of one stream are serialised, also under programmatic dependent
// launch: `pdl_wait()` returns only when the previous grid has completed), so split-mode convolutions may run concurrently
// on different streams (pipeline.SparseInference.run_many drives one stream per host thread).
#include <mutex>
namespace scn {
struct SplitWs {
    cudaStream_t stream;
    float* scratch;
    size_t bytes;
    unsigned int* tickets;
    int ntickets;
    bool used;
};
static SplitWs g_ws[16];
static std::mutex g_ws_mutex;
int split_workspace(cudaStream_t stream, size_t bytes, int n_tiles, float** scratch, unsigned int** tickets) {
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    SplitWs* w = nullptr;
    for (auto& e : g_ws)
        if (e.used && e.stream == stream) w = &e;
    if (!w)
        for (auto& e : g_ws)
            if (!e.used) {
                w = &e;
                *w = SplitWs{stream, nullptr, 0, nullptr, 0, true};
                break;
            }
    if (!w) {
        set_error("conv_fwd_tf32: split-mode convolutions on more than 16 streams");
        return SCN_ERR_INVALID;
    }
    if (bytes > w->bytes) {
        if (w->scratch) {
            cudaDeviceSynchronize();
            cudaFree(w->scratch);
        }
        w->scratch = nullptr, w->bytes = 0;
        size_t want = bytes + bytes / 4;
        if (cudaMalloc(&w->scratch, want) != cudaSuccess || cudaMemsetAsync(w->scratch, 0, want, stream) != cudaSuccess) {
            cudaGetLastError();
            set_error("conv_fwd_tf32: split workspace of %zu bytes: allocation failed", want);
            return SCN_ERR_CUDA;
        }
        w->bytes = want;
    }
    if (n_tiles > w->ntickets) {
        if (w->tickets) {
            cudaDeviceSynchronize();
            cudaFree(w->tickets);
        }
        w->tickets = nullptr, w->ntickets = 0;
        int want = n_tiles < 1024 ? 1024 : 2 * n_tiles;
        if (cudaMalloc(&w->tickets, want * sizeof(unsigned int)) != cudaSuccess ||
            cudaMemsetAsync(w->tickets, 0, want * sizeof(unsigned int), stream) != cudaSuccess) {
            cudaGetLastError();
            set_error("conv_fwd_tf32: ticket allocation failed");
            return SCN_ERR_CUDA;
        }
        w->ntickets = want;
    }
    *scratch = w->scratch, *tickets = w->tickets;
    return SCN_OK;
}
}  // namespace scn

// cuTensorMapEncodeTiled is resolved through the runtime (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        cudaDriverEntryPointQueryResult q;
        void* ptr = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
        else
            cudaGetLastError();
    }
    return fn;
}
// 2-D map over a row-major fp32 feature tensor [rows, C] for tile::gather4: box = 32 columns x 1 row,
// TFLOAT32 (round-to-nearest conversion on load), out-of-bounds -> zeros.
int scn::make_gather_tmap(CUtensorMap* tm, const float* base, int rows, int C, int ld, CUtensorMapSwizzle swz) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return SCN_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)(rows > 0 ? rows : 1)};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32u, 1u};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%d C=%d ld=%d", (int)rc, rows, C, ld);
        return SCN_ERR_CUDA;
    }
    return SCN_OK;
}

static inline int pad16(int c) { return (c + 15) / 16 * 16; }
static inline int n_kblocks(int cin) { return (cin + KB - 1) / KB; }

extern "C" {

int64_t scn_conv_weight_image_bytes(int K, int Cin, int Cout) {
    return (int64_t)K * n_kblocks(Cin) * pad16(Cout) * 128;
}

int scn_conv_pack_weights(const float* w, int K, int Cin, int Cout, int transpose, int reverse, void* image,
                          scn_stream_t stream) {
    SCN_REQUIRE(K > 0 && Cin > 0 && Cout > 0, "pack_weights: bad shape");
    int A = transpose ? Cout : Cin, B = transpose ? Cin : Cout;
    int64_t total = scn_conv_weight_image_bytes(K, Cin, Cout) / 4;
    k_pack_weights<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(w, K, A, B, transpose, reverse, Cin, Cout, pad16(Cout),
                                                                        n_kblocks(Cin), reinterpret_cast<float*>(image));
    return check_launch("pack_weights");
}

#ifdef SCN_EXP_TRACE
int scn_debug_read_trace(unsigned long long* out) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, scn::g_trace, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : 1;
}
int scn_debug_read_cta_times(unsigned long long* out) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, scn::g_cta_times, sizeof(unsigned long long) * 1024) == cudaSuccess ? 0 : 1;
}
#endif
int scn_conv_pack_weights_multi(const int64_t* table, int n, scn_stream_t stream) {
    SCN_REQUIRE(n >= 0 && (n == 0 || table), "pack_weights_multi: bad table");
    if (n == 0) return SCN_OK;
    // 24 blocks x 256 threads per image: the largest image of the shipped networks (27 x 112 x 112) is 1.4 M elements
    dim3 grid(24, n);
    k_pack_weights_multi<<<grid, 256, 0, as_stream(stream)>>>(table);
    return check_launch("pack_weights_multi");
}

int scn_conv_fwd_tf32(const float* in, int ld_in, int Cin, int n_in, const int32_t* map, int n_out, int K, const void* image,
                      const float* bias, const float* residual, int ld_res, const float* mask, int ld_mask, float* out,
                      int ld_out, int Cout, int epi_flags, scn_stream_t stream) {
    return scn_conv_fwd_tf32_dual(in, ld_in, Cin, n_in, map, n_out, K, image, bias, residual, ld_res, mask, ld_mask, out, ld_out, Cout,
                                  epi_flags, nullptr, 0, 0, stream);
}

int scn_conv_fwd_tf32_dual(const float* in, int ld_in, int Cin, int n_in, const int32_t* map, int n_out, int K, const void* image,
                           const float* bias, const float* residual, int ld_res, const float* mask, int ld_mask, float* out,
                           int ld_out, int Cout, int epi_flags, float* out2, int ld_out2, int epi2_flags, scn_stream_t stream) {
    SCN_REQUIRE(!out2 || (ld_out2 >= Cout && !(epi2_flags & ~(SCN_EPI_RELU | SCN_EPI_ROUND)) && out2 != out),
                "conv_fwd_tf32: second output takes RELU / ROUND only and its own buffer");
    SCN_REQUIRE(Cin > 0 && Cout > 0 && K > 0, "conv_fwd_tf32: bad shape Cin=%d Cout=%d K=%d", Cin, Cout, K);
    SCN_REQUIRE(Cout <= 256, "conv_fwd_tf32: Cout > 256 not supported (got %d)", Cout);
    SCN_REQUIRE(map || K == 1, "conv_fwd_tf32: identity map requires K == 1");
    SCN_REQUIRE(!(epi_flags & SCN_EPI_ADD) || residual, "conv_fwd_tf32: SCN_EPI_ADD needs a residual pointer");
    SCN_REQUIRE(!(epi_flags & SCN_EPI_MASK) || mask, "conv_fwd_tf32: SCN_EPI_MASK needs a mask pointer");
    SCN_REQUIRE((reinterpret_cast<uintptr_t>(image) & 15) == 0, "conv_fwd_tf32: weight image must be 16-byte aligned");
    SCN_REQUIRE((reinterpret_cast<uintptr_t>(in) & 3) == 0, "conv_fwd_tf32: input not 4-byte aligned");
    if (n_out <= 0) return SCN_OK;
    static int is100 = -1;
    if (is100 < 0) is100 = scn_device_is_sm100();
    SCN_REQUIRE(is100 == 1, "conv_fwd_tf32: needs an sm_100 device (tcgen05)");
    {
        // submanifold 3^3 layers over a map with an attached tile book (Morton-ordered rows): tile-local kernel, conv_ts.cu
        const int ts = scn::conv_ts_try(in, ld_in, Cin, map, n_out, K, image, bias, residual, ld_res, mask, ld_mask, out, ld_out,
                                        Cout, epi_flags, out2, ld_out2, epi2_flags, as_stream(stream));
        if (ts > 0) return SCN_OK;
        if (ts < 0) return -ts;
    }

    ConvTcParams p;
    p.in = in, p.ld_in = ld_in, p.Cin = Cin, p.map = map, p.n_out = n_out, p.K = K;
    p.image = reinterpret_cast<const uint8_t*>(image), p.bias = bias, p.residual = residual, p.ld_res = ld_res;
    p.mask = mask, p.ld_mask = ld_mask;
    p.out = out, p.ld_out = ld_out, p.Cout = Cout, p.epi = epi_flags;
    p.out2 = out2, p.ld_out2 = ld_out2, p.epi2 = epi2_flags;
    p.cout_pad = pad16(Cout), p.n_kb = n_kblocks(Cin), p.cin_pad8 = (Cin + 7) / 8 * 8;
    p.n_tiles = cdiv(n_out, TILE_M);
    int cols = 2 * p.cout_pad, tc = 32;
    while (tc < cols) tc <<= 1;
    p.tmem_cols = tc;
    const int stage_bytes = A_STAGE_BYTES + p.cout_pad * 128;
    // ring depth = units in flight per CTA (the gather is latency bound, so deeper is better): two CTAs
    // per SM when at least four stages fit in ~110 KB each, otherwise one CTA with up to eight stages
    // The kernel is bound by a per-CTA latency chain (profiles/r1_d_producer_trace.md: identical per-CTA unit period
    // at one and two CTAs per SM), so residency beats ring depth: prefer three CTAs per SM with >= 3 stages, then two
    // CTAs with >= 3 stages, else one CTA with up to eight stages.
    int stages, ctas_per_sm;
    static int max_ctas = -1;
    if (max_ctas < 0) {
        const char* e = getenv("SCN_CONV_CTAS");
        max_ctas = e ? atoi(e) : 3;
    }
    // candidates: (CTAs per SM, smem budget per CTA); every CTA needs >= 3 stages (>= 2 for the 4-CTA case is not enough
    // to overlap gather, MMA and hand-over) and its TMEM columns must fit 512 / CTAs
    stages = 0, ctas_per_sm = 1;
    const int budgets[4] = {220 * 1024, 110 * 1024, 72 * 1024, 54 * 1024};
    for (int c = (max_ctas > 4 ? 4 : max_ctas); c >= 1 && !stages; --c) {
        const int st = budgets[c - 1] / stage_bytes;
        if ((st >= 3 || c == 1) && c * tc <= 512) stages = st > 8 ? 8 : st, ctas_per_sm = c;
    }
    SCN_REQUIRE(stages >= 2, "conv_fwd_tf32: tile does not fit in shared memory");
    p.stages = stages;
    const int smem = stages * stage_bytes + 1024 + 256;
    int vec = 1;
    if (Cin % 4 == 0 && ld_in % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) vec = 4;
    else if (Cin % 2 == 0 && ld_in % 2 == 0 && (reinterpret_cast<uintptr_t>(in) & 7) == 0) vec = 2;
    // TMA gather needs a 16-byte aligned base and row stride
    // Measured on B200 (profiles/r1_c_tma_gather4.md): the TMA unit retires ~one 128-byte row per 14 cycles per
    // SM and spends the same on zero-filled (inactive) rows, so tile::gather4 is 2.2x SLOWER than the cp.async
    // producers for this access pattern.  It stays available as an opt-in (SCN_CONV_TMA=1).
    static int want_tma = -1;
    if (want_tma < 0) {
        const char* e = getenv("SCN_CONV_TMA");
        want_tma = (e && e[0] == '1') ? 1 : 0;
    }
    const bool use_tma = want_tma && (ld_in % 4 == 0) && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && n_in > 0;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (use_tma) {
        int rc = scn::make_gather_tmap(&tmap, in, n_in, Cin, ld_in, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    // small levels: one CTA would walk all K * n_kb units of its tile serially (~0.6 us each) while most SMs idle
    // (profiles/r1_e_launches_fused.md: 40-60 us per layer at 1..122 tiles).  Split the offsets of a tile over
    // several work items; partial sums are accumulated with fp32 atomics into the bias-prefilled output.
    const int slots = sm_count() * ctas_per_sm;
    p.osplit = 1, p.opg = K;
    static int no_split = -1;
    if (no_split < 0) {
        const char* e = getenv("SCN_CONV_NOSPLIT");
        no_split = (e && e[0] == '1') ? 1 : 0;
    }
    // split mode as a thread-block cluster (<= 8 CTAs, the portable limit) reducing through DSMEM when the partial tile
    // (128 rows x (cout_pad + 4) floats) fits in the stage ring; otherwise atomics into the library workspace
    const char* ev_nc = getenv("SCN_CONV_NOCLUSTER");      // read per call: the fallback path is exercised by a test
    const int no_cluster = (ev_nc && ev_nc[0] == '1') ? 1 : 0;
    const bool cluster_ok = !no_cluster && !use_tma && TILE_M * (p.cout_pad * 4 + 16) <= stages * stage_bytes;
    if (!no_split && K > 1 && p.n_tiles * 2 <= slots) {
        int g = slots / p.n_tiles;
        if (g > K) g = K;
        if (g > 8) g = 8;      // one cluster per tile (portable cluster size limit)
        p.opg = cdiv(K, g);
        p.osplit = cdiv(K, p.opg);
    }
    // (Measured and dropped, profiles/r1_m_row_skipping.md: the same cluster split for mid-size levels -- 0.5 .. 3 tiles per
    // slot, where the persistent grid's last wave is mostly empty -- with one work item per CTA and a grid larger than the
    // resident slots.  A non-persistent work item pays its prologue and the cluster reduction every time: level 1 of the
    // bench scene 82.6 -> 82.9 / 85.6 / 90.4 us at cluster size 2 / 3 / 4.)
    p.n_whole = p.osplit > 1 ? 0 : p.n_tiles;
    // Tail split: a persistent grid walks whole tiles round robin, so a level with slots < tiles costs ceil(tiles / slots)
    // tile times even when the last wave is nearly empty (level 1 of the bench scene: 503 tiles on 444 slots = two tile
    // times, 83 us against 52 us for level 0 with 2.6x the rows; profiles/r1_m_row_skipping.md).  The `rem` tiles of the
    // last wave are split into offset groups instead, one group per otherwise idle CTA, and use the atomics + last-arriver
    // epilogue of split mode (the cluster reduction needs one work item per CTA).  SCN_CONV_TAILSPLIT=0/1 overrides.
    {
        const char* ev_ts = getenv("SCN_CONV_TAILSPLIT");
        const int tail_on = ev_ts ? (ev_ts[0] == '1') : SCN_CONV_TAILSPLIT_DEFAULT;
        const int rem = p.n_tiles % slots;
        if (tail_on && !no_split && !use_tma && K > 1 && p.osplit == 1 && p.n_tiles > slots && rem > 0 && rem * 2 <= slots) {
            int g = slots / rem;
            if (g > K) g = K;
            if (g > 8) g = 8;      // depth of the same-address reduction chain
            p.opg = cdiv(K, g);
            p.osplit = cdiv(K, p.opg);
            if (p.osplit > 1) p.n_whole = p.n_tiles - rem;
            else p.opg = K;
        }
    }
    const int n_work = p.n_whole + (p.n_tiles - p.n_whole) * p.osplit;
    p.n_work = n_work;
    int grid = n_work < slots ? n_work : slots;
    p.scratch = nullptr, p.tickets = nullptr, p.cluster = 0;
    p.skip = scn::conv_row_skipping();
    if (p.osplit > 1 && p.n_whole == 0 && p.osplit <= 8 && cluster_ok && grid == n_work) p.cluster = p.osplit;
    if (p.osplit > 1 && !p.cluster) {
        int rc = scn::split_workspace(as_stream(stream), (size_t)n_out * Cout * sizeof(float), p.n_tiles, &p.scratch, &p.tickets);
        if (rc) return rc;
    }
    cudaError_t e;
    auto launch = [&](auto kern, int threads) {
        e = (cudaError_t)scn::ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem);
        if (e != cudaSuccess) return;
        scn::PdlLaunch L(dim3(grid), dim3(threads), smem, as_stream(stream), p.cluster);
        e = cudaLaunchKernelEx(&L.cfg, kern, tmap, p);
    };
    if (use_tma) launch(k_conv_tc<4, true>, 192);
    else if (vec == 4) launch(k_conv_tc<4, false>, CONV_THREADS);
    else if (vec == 2) launch(k_conv_tc<2, false>, CONV_THREADS);
    else launch(k_conv_tc<1, false>, CONV_THREADS);
    if (e != cudaSuccess) {
        cudaGetLastError();
        scn::set_error("conv_fwd_tf32: cudaFuncSetAttribute(%d bytes): %s", smem, cudaGetErrorString(e));
        return SCN_ERR_CUDA;
    }
    return check_launch("conv_fwd_tf32");
}

}  // extern "C"
