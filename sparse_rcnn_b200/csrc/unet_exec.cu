// Native executor of the sparse U-Net (encoder pyramid + decoder with skip connections) that the reference assembles in
// ndsis/modules/module_factory.py:438-578,789-830 and runs through model.py:414-446 (FeatureExtractor.forward):
//
//   encoder level i : [SubM K^3 | Convolution 2^3/s2]  ->  num_units residual units            E_i
//   decoder level j : ReLU -> Deconvolution 2^3/s2 -> JoinTable(up, skip E_l) -> NetworkInNetwork -> residual units   D_j
//
// One C-ABI call per direction walks a LAYER TABLE (static: channel counts, parameter / packed-image pointers) against a
// GEOMETRY TABLE (per scene: rows per level, neighbour maps) and enqueues every kernel of the pass on the caller's stream;
// all activations live in ONE caller-allocated arena whose layout `scn_unet_plan` computes from the two tables.  Why: the
// training step was host bound (profiles/r1_i_host_bound.md: 195 Python->C calls, 305 allocations, 47 autograd nodes per
// step for ~380 kernels); the executor turns the backbone into 2 calls, 2 allocations and 1 autograd node without changing
// which kernels run or in which order -- results are bit-identical to the per-layer entry points it sequences
// (scn_conv_layer_*, scn_residual_unit_*; tests/test_gpu_executor.py).
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include "common.cuh"

using namespace scn;

#define SCN_TRY(call)          \
    do {                       \
        int rc__ = (call);     \
        if (rc__) return rc__; \
    } while (0)

namespace {

constexpr int MAX_LEVELS = 8, MAX_UNITS = 4;
#ifndef SCN_EXEC_BWD_PDL_DEFAULT
#define SCN_EXEC_BWD_PDL_DEFAULT 0
#endif

struct Conv {      // one convolution layer; kind 0 = absent (the level passes its input through)
    int kind, K, cin, cout;
    const float *w, *b;
    void *img_f, *img_b;
};
struct Unit {
    const float *w1, *b1, *w2, *b2;
    void *i1f, *i2f, *i1b, *i2b;
};
struct Net {
    int L;                                    // encoder levels; decoder levels = L - 1
    Conv enc[MAX_LEVELS];
    int enc_units[MAX_LEVELS];
    Unit eu[MAX_LEVELS][MAX_UNITS];
    Conv deconv[MAX_LEVELS], nin[MAX_LEVELS];      // decoder level j writes level l = L - 2 - j
    int dec_units[MAX_LEVELS];
    Unit du[MAX_LEVELS][MAX_UNITS];
    int C[MAX_LEVELS];                        // channels of E_i
    int CD[MAX_LEVELS];                       // channels of D_j
};
struct Geo {
    int n[MAX_LEVELS];
    const int32_t* subm[MAX_LEVELS];          // [27, n_i]
    const int32_t* cmap[MAX_LEVELS];          // level i -> i+1: [8, n_{i+1}] children of each coarse row
    const int32_t* dmap[MAX_LEVELS];          // level i+1 -> i: [8, n_i]   parent of each fine row per in-box offset
};

// ---- table formats (int64 words), written by sparse_rcnn_b200/executor.py -------------------------------------------------
//  net : [L] then per encoder level  [kind K cin cout w b img_f img_b n_units] + n_units x [w1 b1 w2 b2 i1f i2f i1b i2b]
//        then per decoder level      [K cin cout w b img_f img_b] (deconvolution) [cin cout w b img_f img_b] (1x1 layer)
//                                    [n_units] + n_units x unit record
//  geo : per level [n subm_map cmap_to_next dmap_from_next]
template <class T>
static T* as_ptr(int64_t v) { return reinterpret_cast<T*>(static_cast<uintptr_t>(v)); }

static const int64_t* read_unit(const int64_t* p, Unit& u) {
    u.w1 = as_ptr<const float>(p[0]), u.b1 = as_ptr<const float>(p[1]), u.w2 = as_ptr<const float>(p[2]), u.b2 = as_ptr<const float>(p[3]);
    u.i1f = as_ptr<void>(p[4]), u.i2f = as_ptr<void>(p[5]), u.i1b = as_ptr<void>(p[6]), u.i2b = as_ptr<void>(p[7]);
    return p + 8;
}

static int parse_net(const int64_t* p, Net& net) {
    std::memset(&net, 0, sizeof(net));
    net.L = (int)*p++;
    SCN_REQUIRE(net.L >= 1 && net.L <= MAX_LEVELS, "unet: %d levels (1..%d supported)", net.L, MAX_LEVELS);
    for (int i = 0; i < net.L; ++i) {
        Conv& c = net.enc[i];
        c.kind = (int)p[0], c.K = (int)p[1], c.cin = (int)p[2], c.cout = (int)p[3];
        c.w = as_ptr<const float>(p[4]), c.b = as_ptr<const float>(p[5]), c.img_f = as_ptr<void>(p[6]), c.img_b = as_ptr<void>(p[7]);
        net.enc_units[i] = (int)p[8];
        p += 9;
        SCN_REQUIRE(net.enc_units[i] >= 0 && net.enc_units[i] <= MAX_UNITS, "unet: %d units per level (max %d)", net.enc_units[i], MAX_UNITS);
        SCN_REQUIRE(c.kind == 0 ? (i == 0 && net.enc_units[i] == 0) : (c.kind == 1 ? (c.K == 1 || c.K == 27) : (c.kind == 2 && c.K == 8 && i > 0)),
                    "unet: bad entry layer at level %d", i);
        for (int u = 0; u < net.enc_units[i]; ++u) p = read_unit(p, net.eu[i][u]);
        net.C[i] = c.cout;
    }
    for (int j = 0; j < net.L - 1; ++j) {
        Conv& d = net.deconv[j];
        d.kind = 3, d.K = (int)p[0], d.cin = (int)p[1], d.cout = (int)p[2];
        d.w = as_ptr<const float>(p[3]), d.b = as_ptr<const float>(p[4]), d.img_f = as_ptr<void>(p[5]), d.img_b = as_ptr<void>(p[6]);
        p += 7;
        Conv& m = net.nin[j];
        m.kind = 4, m.K = 1, m.cin = (int)p[0], m.cout = (int)p[1];
        m.w = as_ptr<const float>(p[2]), m.b = as_ptr<const float>(p[3]), m.img_f = as_ptr<void>(p[4]), m.img_b = as_ptr<void>(p[5]);
        p += 6;
        net.dec_units[j] = (int)*p++;
        SCN_REQUIRE(net.dec_units[j] >= 0 && net.dec_units[j] <= MAX_UNITS, "unet: %d units per level (max %d)", net.dec_units[j], MAX_UNITS);
        for (int u = 0; u < net.dec_units[j]; ++u) p = read_unit(p, net.du[j][u]);
        const int l = net.L - 2 - j;
        SCN_REQUIRE(d.K == 8 && d.cin == (j == 0 ? net.C[net.L - 1] : net.CD[j - 1]) && m.cin == d.cout + net.C[l],
                    "unet: decoder level %d does not fit the encoder (channels)", j);
        net.CD[j] = m.cout;
    }
    return SCN_OK;
}

static void parse_geo(const int64_t* p, const Net& net, Geo& g) {
    std::memset(&g, 0, sizeof(g));
    for (int i = 0; i < net.L; ++i, p += 4) {
        g.n[i] = (int)p[0];
        g.subm[i] = as_ptr<const int32_t>(p[1]), g.cmap[i] = as_ptr<const int32_t>(p[2]), g.dmap[i] = as_ptr<const int32_t>(p[3]);
    }
}

// ---- arena layout -----------------------------------------------------------------------------------------------------------
struct Cursor {
    int64_t off = 0;      // in floats; every buffer starts on a 256-byte boundary
    int64_t take(int64_t rows, int64_t cols) {
        const int64_t at = off;
        off += (rows * cols + 63) / 64 * 64;
        return at;
    }
};
struct Plan {
    // forward (kept for the backward)
    int64_t xr[MAX_LEVELS], c[MAX_LEVELS], er[MAX_LEVELS][MAX_UNITS], eh[MAX_LEVELS][MAX_UNITS], ey[MAX_LEVELS][MAX_UNITS];
    int64_t rl[MAX_LEVELS], cat[MAX_LEVELS], catr[MAX_LEVELS], nin[MAX_LEVELS];
    int64_t dr[MAX_LEVELS][MAX_UNITS], dh[MAX_LEVELS][MAX_UNITS], dy[MAX_LEVELS][MAX_UNITS];
    int64_t E[MAX_LEVELS], D[MAX_LEVELS];      // outputs (E[0] = -1 when level 0 passes the input through)
    int64_t fwd_total;
    // backward
    int64_t gD[MAX_LEVELS], gE[MAX_LEVELS], tmp[MAX_LEVELS];      // gradient wrt D_j / E_i, a level-sized temporary
    // scratch of every unit and layer is its own: the weight gradients run on a side stream and read these buffers while
    // the main stream is already working on the next layer
    int64_t e_gyr[MAX_LEVELS][MAX_UNITS], e_gh[MAX_LEVELS][MAX_UNITS], e_gx[MAX_LEVELS][MAX_UNITS], e_round[MAX_LEVELS];
    int64_t d_gyr[MAX_LEVELS][MAX_UNITS], d_gh[MAX_LEVELS][MAX_UNITS], d_gx[MAX_LEVELS][MAX_UNITS], d_round_nin[MAX_LEVELS],
        d_round_up[MAX_LEVELS];
    int64_t g_cat[MAX_LEVELS], g_up[MAX_LEVELS], g_skip[MAX_LEVELS], g_rl[MAX_LEVELS], g_catr[MAX_LEVELS];
    int64_t bwd_total;
};

static void make_plan(const Net& net, const Geo& g, Plan& P) {
    std::memset(&P, 0, sizeof(P));
    Cursor f;
    for (int i = 0; i < net.L; ++i) {
        const Conv& e = net.enc[i];
        const int n = g.n[i], n_in = i == 0 ? g.n[0] : g.n[i - 1];
        if (e.kind) {
            P.xr[i] = f.take(n_in, e.cin);
            P.c[i] = f.take(n, e.cout);
        }
        for (int u = 0; u < net.enc_units[i]; ++u)
            P.er[i][u] = f.take(n, net.C[i]), P.eh[i][u] = f.take(n, net.C[i]), P.ey[i][u] = f.take(n, net.C[i]);
        P.E[i] = net.enc_units[i] ? P.ey[i][net.enc_units[i] - 1] : (e.kind ? P.c[i] : -1);
    }
    for (int j = 0; j < net.L - 1; ++j) {
        const int l = net.L - 2 - j, n = g.n[l], c = net.CD[j], cw = net.nin[j].cin;
        P.rl[j] = f.take(g.n[l + 1], net.deconv[j].cin);
        P.cat[j] = f.take(n, cw), P.catr[j] = f.take(n, cw), P.nin[j] = f.take(n, c);
        for (int u = 0; u < net.dec_units[j]; ++u) P.dr[j][u] = f.take(n, c), P.dh[j][u] = f.take(n, c), P.dy[j][u] = f.take(n, c);
        P.D[j] = net.dec_units[j] ? P.dy[j][net.dec_units[j] - 1] : P.nin[j];
    }
    P.fwd_total = f.off;
    Cursor b;
    for (int i = 0; i < net.L; ++i) {
        int cmax = net.C[i];
        for (int j = 0; j < net.L - 1; ++j)
            if (net.L - 2 - j == i && net.nin[j].cin > cmax) cmax = net.nin[j].cin;
        const int n = g.n[i];
        P.gE[i] = b.take(n, net.C[i]), P.tmp[i] = b.take(n, cmax);
        P.e_round[i] = b.take(n, net.C[i]);
        for (int u = 0; u < net.enc_units[i]; ++u)
            P.e_gyr[i][u] = b.take(n, net.C[i]), P.e_gh[i][u] = b.take(n, net.C[i]), P.e_gx[i][u] = b.take(n, net.C[i]);
    }
    for (int j = 0; j < net.L - 1; ++j) {
        const int l = net.L - 2 - j, n = g.n[l];
        P.gD[j] = b.take(n, net.CD[j]);
        P.d_round_nin[j] = b.take(n, net.CD[j]), P.d_round_up[j] = b.take(n, net.deconv[j].cout);
        for (int u = 0; u < net.dec_units[j]; ++u)
            P.d_gyr[j][u] = b.take(n, net.CD[j]), P.d_gh[j][u] = b.take(n, net.CD[j]), P.d_gx[j][u] = b.take(n, net.CD[j]);
        P.g_cat[j] = b.take(n, net.nin[j].cin), P.g_up[j] = b.take(n, net.deconv[j].cout), P.g_skip[j] = b.take(n, net.C[l]);
        P.g_rl[j] = b.take(g.n[l + 1], net.deconv[j].cin);
        P.g_catr[j] = b.take(n, net.nin[j].cin);
    }
    P.bwd_total = b.off;
}

// dst[r][0:cols] = src[r][0:cols] (+ dst2 = round-to-nearest TF32 of the same values): the column blocks of JoinTable and its
// backward.  Own kernel instead of cudaMemcpy2DAsync: the 2-D copy engine path moved the 21 MB of a level-0 skip connection
// in 34 us, 16-byte vector accesses in a grid-stride loop take ~12 us, and the optional second output replaces the rounding
// pass over the joined buffer.
template <bool DUAL>
__global__ void __launch_bounds__(256) k_copy_cols(float* __restrict__ dst, int ld_dst, float* __restrict__ dst2, int ld_dst2,
                                                   const float* __restrict__ src, int ld_src, int64_t total, int q) {
    pdl_trigger();      // PDL (common.cuh): the convolution that follows sets itself up meanwhile
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t r = i / q;
        const int c = (int)(i - r * q) * 4;
        const float4 v = *reinterpret_cast<const float4*>(src + r * ld_src + c);
        *reinterpret_cast<float4*>(dst + r * ld_dst + c) = v;
        if (DUAL) *reinterpret_cast<float4*>(dst2 + r * ld_dst2 + c) = epi2_apply4(v, SCN_EPI_ROUND);
    }
}

static int copy_cols(float* dst, int ld_dst, const float* src, int ld_src, int n, int cols, cudaStream_t st, float* dst2 = nullptr,
                     int ld_dst2 = 0) {
    if (n <= 0 || cols <= 0) return SCN_OK;
    const uintptr_t al = reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst2);
    if ((cols | ld_dst | ld_src | ld_dst2) % 4 == 0 && (al & 15) == 0) {
        const int q = cols / 4;
        const int64_t total = (int64_t)n * q;
        const int grid = grid_for(total, 256);
        // programmatic dependent launch for this kernel: measured on one box, three alternations -- forward 1.478-1.479 ms with
        // the attribute, 1.464-1.466 ms without (a 1184-CTA elementwise grid parked in griddepcontrol.wait takes the SMs the
        // small-level convolution in front of it is still using); SCN_COPY_PDL=1 turns it on
        static int copy_pdl = -1;
        if (copy_pdl < 0) {
            const char* e = getenv("SCN_COPY_PDL");
            copy_pdl = (e && e[0] == '1') ? 1 : 0;
        }
        PdlMaskScope scope;
        if (!copy_pdl) scope.exclude(st);
        PdlLaunch L(dim3(grid), dim3(256), 0, st);
        float* no2 = nullptr;
        const int zero = 0;
        if (dst2) cudaLaunchKernelEx(&L.cfg, k_copy_cols<true>, dst, ld_dst, dst2, ld_dst2, src, ld_src, total, q);
        else cudaLaunchKernelEx(&L.cfg, k_copy_cols<false>, dst, ld_dst, no2, zero, src, ld_src, total, q);
        return check_launch("unet: copy_cols");
    }
    SCN_REQUIRE(!dst2, "unet: copy_cols with a second output needs 16-byte aligned column blocks");
    cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)ld_dst * 4, src, (size_t)ld_src * 4, (size_t)cols * 4, (size_t)n,
                                      cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) {
        set_error("unet: cudaMemcpy2DAsync: %s", cudaGetErrorString(e));
        return SCN_ERR_CUDA;
    }
    return SCN_OK;
}

// Second outputs (scn_conv_fwd_tf32_dual): in TF32 mode the convolution that produces a tensor also writes the operand its
// consumer gathers -- relu + round for the next residual unit / the decoder's ReLU -> Deconvolution, round for the next
// level's strided convolution, round of a gradient for the next transposed convolution -- instead of one elementwise
// kernel per consumer (54 of the ~380 launches of a training step; every one of them sat on the dependency chain).  Same
// values as the separate passes, bit for bit.  SCN_EXEC_DUAL=0 restores the separate kernels (read per call: tests).
static bool dual_outputs() {
    const char* e = getenv("SCN_EXEC_DUAL");
    return !(e && e[0] == '0');
}
struct Out2 {
    float* p;
    int epi;
};
static const Out2 NO_OUT2 = {nullptr, 0};

// a chain of residual units over one level, forward.  r_ready: r[0] already holds relu(x) (rounded), written by x's producer;
// tail: second output of the chain's last convolution
static int stage_fwd(const Unit* units, int U, const float* x, bool r_ready, int n, int C, const int32_t* map, float* base,
                     const int64_t* r, const int64_t* h, const int64_t* y, bool dual, Out2 tail, int tf32, scn_stream_t s) {
    for (int u = 0; u < U; ++u) {
        const Unit& q = units[u];
        if (!tf32) {
            SCN_TRY(scn_residual_unit_fwd(x, n, C, map, 27, q.w1, q.b1, q.w2, q.b2, q.i1f, q.i2f, 0, base + r[u], base + h[u], base + y[u],
                                          tf32, s));
        } else if (n > 0) {
            if (!(dual && (u > 0 || r_ready))) SCN_TRY(scn_relu_fwd(x, base + r[u], (int64_t)n * C, 1, s));
            SCN_TRY(scn_conv_fwd_tf32(base + r[u], C, C, n, map, n, 27, q.i1f, q.b1, nullptr, 0, nullptr, 0, base + h[u], C, C,
                                      SCN_EPI_RELU | SCN_EPI_ROUND, s));
            const Out2 o = !dual ? NO_OUT2 : (u + 1 < U ? Out2{base + r[u + 1], SCN_EPI_RELU | SCN_EPI_ROUND} : tail);
            SCN_TRY(scn_conv_fwd_tf32_dual(base + h[u], C, C, n, map, n, 27, q.i2f, q.b2, x, C, nullptr, 0, base + y[u], C, C, SCN_EPI_ADD,
                                           o.p, C, o.epi, s));
        }
        x = base + y[u];
    }
    return SCN_OK;
}

// ---- weight gradients on a side stream -----------------------------------------------------------------------------------------
// In the backward pass only the input gradients form a dependency chain; a layer's weight (+ bias) gradient is needed by
// nobody before the optimizer.  They are enqueued on a second stream that forks from the main stream where their operands
// are ready and joins it at the end of the call: on the small levels (a handful of tiles, latency bound) they run beside
// the chain instead of in it; on the large levels the SMs are full either way.  SCN_EXEC_SIDE=0 keeps everything in line.
// SCN_EXEC_SIDE_STREAMS=n (default 1) takes n side streams round robin.  Measured and left off: most weight-gradient launches
// belong to the small levels (44 of 68 per step, a handful of CTAs each), so several streams should overlap their latency
// chains -- but the backward call stays at 3.90 / 4.07 / 3.97 / 3.92 ms with 1 / 2 / 3 / 4 side streams (profiles/r2_h): the
// backward is bound by the two big levels' throughput, not by the small levels' serialisation.
constexpr int MAX_SIDE = 4;
struct Side {
    cudaStream_t main, sides[MAX_SIDE];
    int n_side = 0;
    bool on;
    cudaEvent_t ev[16];
    int next = 0, cur = 0;
    unsigned used = 0;      // bit i: sides[i] has work of this call
    int fork() {      // the next side stream waits for everything enqueued on the main stream so far
        if (!on) return SCN_OK;
        cur = (cur + 1) % n_side;
        cudaEvent_t e = ev[next++ & 15];
        if (cudaEventRecord(e, main) != cudaSuccess || cudaStreamWaitEvent(sides[cur], e, 0) != cudaSuccess) {
            set_error("unet_bwd: stream fork: %s", cudaGetErrorString(cudaGetLastError()));
            return SCN_ERR_CUDA;
        }
        used |= 1u << cur;
        return SCN_OK;
    }
    int join() {
        if (!on) return SCN_OK;
        for (int i = 0; i < n_side; ++i) {
            if (!(used >> i & 1u)) continue;
            cudaEvent_t e = ev[next++ & 15];
            if (cudaEventRecord(e, sides[i]) != cudaSuccess || cudaStreamWaitEvent(main, e, 0) != cudaSuccess) {
                set_error("unet_bwd: stream join: %s", cudaGetErrorString(cudaGetLastError()));
                return SCN_ERR_CUDA;
            }
        }
        used = 0;
        return SCN_OK;
    }
    scn_stream_t wstream() const { return reinterpret_cast<scn_stream_t>(on ? sides[cur] : main); }
};

struct SideResources {
    cudaStream_t sides[MAX_SIDE];
    cudaEvent_t ev[16];
};
static int side_for(cudaStream_t main, Side& S) {
    static std::mutex mu;
    static std::unordered_map<cudaStream_t, SideResources> pool;      // side streams per calling stream (host threads)
    static int enabled = -1, n_side = 1;
    std::lock_guard<std::mutex> lock(mu);
    if (enabled < 0) {
        const char* e = getenv("SCN_EXEC_SIDE");
        enabled = (e && e[0] == '0') ? 0 : 1;
        const char* ns = getenv("SCN_EXEC_SIDE_STREAMS");
        if (ns) n_side = atoi(ns);
        n_side = n_side < 1 ? 1 : (n_side > MAX_SIDE ? MAX_SIDE : n_side);
    }
    S.main = main, S.on = false, S.n_side = 0;
    if (!enabled) return SCN_OK;
    auto it = pool.find(main);
    if (it == pool.end()) {
        SideResources r;
        for (auto& st : r.sides)
            if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
                set_error("unet_bwd: cudaStreamCreate: %s", cudaGetErrorString(cudaGetLastError()));
                return SCN_ERR_CUDA;
            }
        for (auto& e : r.ev)
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
                set_error("unet_bwd: cudaEventCreate: %s", cudaGetErrorString(cudaGetLastError()));
                return SCN_ERR_CUDA;
            }
        it = pool.emplace(main, r).first;
    }
    S.on = true, S.n_side = n_side;
    for (int i = 0; i < MAX_SIDE; ++i) S.sides[i] = it->second.sides[i];
    for (int i = 0; i < 16; ++i) S.ev[i] = it->second.ev[i];
    return SCN_OK;
}

// one residual unit backward: the kernels of scn_residual_unit_bwd (accumulate mode), input-gradient chain on the main
// stream, the two weight gradients on the side stream.  gy_state: gyr already holds round(gy) (second output of gy's
// producer) or gy is rounded itself (written by the rounding ReLU backward); gx2: second output of the convolution that
// produces gx
static int unit_bwd(const Unit& q, const float* gy, int gy_state, const float* r, const float* h, int n, int C, const int32_t* map,
                    float* gyr, float* gh, float* gx, Out2 gx2, float* gw1, float* gb1, float* gw2, float* gb2, int tf32, scn_stream_t s,
                    Side& side) {
    if (n == 0) return SCN_OK;
    const int K = 27;
    const int64_t total = (int64_t)n * C;
    const float* g_op = gy;
    if (tf32) {      // gy_state: 0 = round gy into gyr, 1 = gyr already holds round(gy), 2 = gy itself is rounded
        if (gy_state == 0) SCN_TRY(scn_round_tf32(gy, gyr, total, s));
        g_op = gy_state == 2 ? gy : gyr;
    }
    if (gw2 || gb2) {
        SCN_TRY(side.fork());
        if (gw2) SCN_TRY(scn_conv_bwd_weight(h, C, C, map, n, K, g_op, C, C, gw2, gb2, tf32, side.wstream()));
        else SCN_TRY(scn_col_sum_add(gy, C, n, C, gb2, side.wstream()));
    }
    if (tf32)
        SCN_TRY(scn_conv_fwd_tf32(g_op, C, C, n, map, n, K, q.i2b, nullptr, nullptr, 0, h, C, gh, C, C, SCN_EPI_MASK | SCN_EPI_ROUND, s));
    else
        SCN_TRY(scn_conv_fwd_fp32(g_op, C, C, map, n, K, q.w2, 1, 1, nullptr, nullptr, 0, h, C, gh, C, C, SCN_EPI_MASK, s));
    if (gw1 || gb1) {
        SCN_TRY(side.fork());
        if (gw1) SCN_TRY(scn_conv_bwd_weight(r, C, C, map, n, K, gh, C, C, gw1, gb1, tf32, side.wstream()));
        else SCN_TRY(scn_col_sum_add(gh, C, n, C, gb1, side.wstream()));
    }
    if (tf32)
        SCN_TRY(scn_conv_fwd_tf32_dual(gh, C, C, n, map, n, K, q.i1b, nullptr, gy, C, r, C, gx, C, C, SCN_EPI_MASK | SCN_EPI_ADD, gx2.p, C,
                                       gx2.epi, s));
    else
        SCN_TRY(scn_conv_fwd_fp32(gh, C, C, map, n, K, q.w1, 1, 1, nullptr, gy, C, r, C, gx, C, C, SCN_EPI_MASK | SCN_EPI_ADD, s));
    return SCN_OK;
}

// ... a chain of units: gy -> gradient wrt the stage input (returned through *gx_out).  head: where the layer in front of the
// stage wants round(gradient of the stage input) (second output of unit 0's last convolution)
static int stage_bwd(const Unit* units, int U, const float* gy, int n, int C, const int32_t* map, const float* fbase, const int64_t* r,
                     const int64_t* h, float* bbase, const int64_t* gyr, const int64_t* gh, const int64_t* gxs, float* const* pg, bool dual,
                     Out2 head, int gy_state, int tf32, scn_stream_t s, Side& side, const float** gx_out) {
    int ready = gy_state;      // of the chain's incoming gradient (unit_bwd)
    for (int u = U - 1; u >= 0; --u) {
        float* gx = bbase + gxs[u];
        const Out2 o = !dual ? NO_OUT2 : (u > 0 ? Out2{bbase + gyr[u - 1], SCN_EPI_ROUND} : head);
        SCN_TRY(unit_bwd(units[u], gy, ready, fbase + r[u], fbase + h[u], n, C, map, bbase + gyr[u], bbase + gh[u], gx, o, pg[4 * u],
                         pg[4 * u + 1], pg[4 * u + 2], pg[4 * u + 3], tf32, s, side));
        gy = gx;
        ready = (dual && n > 0) ? 1 : 0;
    }
    *gx_out = gy;
    return SCN_OK;
}

// one convolution layer backward: the kernels of scn_conv_layer_bwd, weight gradient on the side stream.  go_exact: go_round
// already holds round(go), both with leading dimension ld_gr; gx2: second output of the input-gradient convolution
static int conv_bwd(const float* go, int n_out, int Cout, bool go_exact, float* go_round, int ld_gr, const float* x, int ld_x, int n_in,
                    int Cin, const int32_t* fmap, const int32_t* bmap, int K, const float* w, void* image_t, int reverse, float* gx, Out2 gx2,
                    float* gw, float* gb, int tf32, scn_stream_t s, Side& side, const float* add = nullptr, const float* mask = nullptr) {
    // add / mask (TF32 mode only): epilogue of the input-gradient convolution -- gx = conv + add (`add` may be gx itself:
    // the skip connection's gradient already sits there) and / or gx = round(mask > 0 ? conv : 0) (the ReLU in front of a
    // transposed convolution), instead of a k_add / k_relu_bwd launch behind it
    if (n_out == 0) {
        if (gx && n_in > 0 && !add) cudaMemsetAsync(gx, 0, sizeof(float) * (size_t)n_in * Cin, as_stream(s));
        if (gx && gx2.p && n_in > 0) {
            if (add) SCN_TRY(scn_round_tf32(add, gx2.p, (int64_t)n_in * Cin, s));
            else cudaMemsetAsync(gx2.p, 0, sizeof(float) * (size_t)n_in * Cin, as_stream(s));
        }
        return check_launch("unet_bwd(memset)");
    }
    const float* g = go;
    int ld_g = Cout;
    if (tf32 && (gx || gw)) {
        if (!go_exact) SCN_TRY(scn_round_tf32(go, go_round, (int64_t)n_out * Cout, s));
        g = go_round, ld_g = go_exact ? ld_gr : Cout;
    }
    if (gw || gb) {
        SCN_TRY(side.fork());
        if (gw) SCN_TRY(scn_conv_bwd_weight(x, ld_x, Cin, fmap, n_out, K, g, ld_g, Cout, gw, gb, tf32, side.wstream()));
        else SCN_TRY(scn_col_sum_add(go, go_exact ? ld_gr : Cout, n_out, Cout, gb, side.wstream()));
    }
    if (gx && n_in > 0) {
        if (tf32)
            SCN_TRY(scn_conv_fwd_tf32_dual(g, ld_g, Cout, n_out, bmap, n_in, K, image_t, nullptr, add, Cin, mask, Cin, gx, Cin, Cin,
                                           (add ? SCN_EPI_ADD : 0) | (mask ? (SCN_EPI_MASK | SCN_EPI_ROUND) : 0), gx2.p, Cin, gx2.epi, s));
        else
            SCN_TRY(scn_conv_fwd_fp32(g, Cout, Cout, bmap, n_in, K, w, 1, reverse, nullptr, nullptr, 0, nullptr, 0, gx, Cin, Cin, 0, s));
    }
    return SCN_OK;
}

}  // namespace

extern "C" {

// offsets (in floats) of the outputs inside the forward arena: out[0 .. L) = E_i (-1: the input itself), out[L .. 2L-1) = D_j;
// out[2L-1] = forward arena size, out[2L] = backward arena size (floats)
int scn_unet_plan(const int64_t* net_table, const int64_t* geo_table, int64_t* out) {
    SCN_REQUIRE(net_table && geo_table && out, "unet_plan: null table");
    Net net;
    SCN_TRY(parse_net(net_table, net));
    Geo g;
    parse_geo(geo_table, net, g);
    Plan P;
    make_plan(net, g, P);
    for (int i = 0; i < net.L; ++i) out[i] = P.E[i];
    for (int j = 0; j < net.L - 1; ++j) out[net.L + j] = P.D[j];
    out[2 * net.L - 1] = P.fwd_total, out[2 * net.L] = P.bwd_total;
    return SCN_OK;
}

int scn_unet_fwd(const int64_t* net_table, const int64_t* geo_table, const float* x, float* arena, int n_decoder_levels,
                 int use_tf32, scn_stream_t stream) {
    SCN_REQUIRE(net_table && geo_table && arena, "unet_fwd: null argument");
    Net net;
    SCN_TRY(parse_net(net_table, net));
    Geo g;
    parse_geo(geo_table, net, g);
    Plan P;
    make_plan(net, g, P);
    cudaStream_t st = as_stream(stream);
    const int tf32 = use_tf32 ? 1 : 0;
    const bool dual = tf32 && dual_outputs();
    const int RR = SCN_EPI_RELU | SCN_EPI_ROUND;
    const int nd = n_decoder_levels < net.L - 1 ? n_decoder_levels : net.L - 1;
    const float* cur = x;      // E_{i-1}
    bool operand_ready = false;      // the consumer's operand of `cur` (xr[i] / rl[0]) was written by cur's producer
    for (int i = 0; i < net.L; ++i) {
        const Conv& e = net.enc[i];
        const int n = g.n[i], n_in = i == 0 ? g.n[0] : g.n[i - 1];
        const int U = net.enc_units[i];
        // what the consumer of E_i gathers: the next level's entry layer reads round(E_i), the first decoder level relu(E_i)
        const Out2 tail = !dual ? NO_OUT2
                          : i + 1 < net.L ? (net.enc[i + 1].kind ? Out2{arena + P.xr[i + 1], SCN_EPI_ROUND} : NO_OUT2)
                                          : (nd > 0 ? Out2{arena + P.rl[0], RR} : NO_OUT2);
        bool r_ready = false;
        if (e.kind) {
            const int32_t* map = e.kind == 2 ? g.cmap[i - 1] : (e.K == 1 ? nullptr : g.subm[i]);
            if (!tf32) {
                SCN_TRY(scn_conv_layer_fwd(cur, e.cin, n_in, e.cin, 0, arena + P.xr[i], map, n, e.K, e.w, e.img_f, 0, e.b, arena + P.c[i],
                                           e.cout, tf32, stream));
            } else if (n > 0) {
                if (!operand_ready) SCN_TRY(scn_round_tf32(cur, arena + P.xr[i], (int64_t)n_in * e.cin, stream));
                const Out2 o = !dual ? NO_OUT2 : (U ? Out2{arena + P.er[i][0], RR} : tail);
                SCN_TRY(scn_conv_fwd_tf32_dual(arena + P.xr[i], e.cin, e.cin, n_in, map, n, e.K, e.img_f, e.b, nullptr, 0, nullptr, 0,
                                               arena + P.c[i], e.cout, e.cout, 0, o.p, e.cout, o.epi, stream));
                r_ready = dual && U > 0;
            }
            cur = arena + P.c[i];
        }
        SCN_TRY(stage_fwd(net.eu[i], U, cur, r_ready, n, net.C[i], g.subm[i], arena, P.er[i], P.eh[i], P.ey[i], dual, tail, tf32, stream));
        if (U) cur = arena + P.E[i];
        operand_ready = dual && tail.p && n > 0 && (e.kind || U);
    }
    for (int j = 0; j < nd; ++j) {
        const int l = net.L - 2 - j, n = g.n[l], n_in = g.n[l + 1];
        const Conv &d = net.deconv[j], &m = net.nin[j];
        const int U = net.dec_units[j];
        // ReLU (rounded in tf32 mode: a valid tensor-core operand) -> transposed convolution over dmap, written into the left
        // columns of the joined buffer; the skip connection is copied beside it (JoinTable)
        if (!operand_ready) SCN_TRY(scn_relu_fwd(cur, arena + P.rl[j], (int64_t)n_in * d.cin, tf32, stream));
        // second outputs of the join: the transposed convolution also writes its columns of catr (rounded), the skip copy
        // writes the other columns of cat and catr -- no rounding pass over the joined buffer
        const bool join2 = dual && (d.cout % 4 == 0) && (m.cin % 4 == 0) && (net.C[l] % 4 == 0);
        if (n > 0) {
            if (tf32)
                SCN_TRY(scn_conv_fwd_tf32_dual(arena + P.rl[j], d.cin, d.cin, n_in, g.dmap[l], n, d.K, d.img_f, d.b, nullptr, 0, nullptr, 0,
                                               arena + P.cat[j], m.cin, d.cout, 0, join2 ? arena + P.catr[j] : nullptr, m.cin,
                                               SCN_EPI_ROUND, stream));
            else
                SCN_TRY(scn_conv_fwd_fp32(arena + P.rl[j], d.cin, d.cin, g.dmap[l], n, d.K, d.w, 0, 0, d.b, nullptr, 0, nullptr, 0,
                                          arena + P.cat[j], m.cin, d.cout, 0, stream));
        }
        const float* skip = P.E[l] >= 0 ? arena + P.E[l] : x;
        SCN_TRY(copy_cols(arena + P.cat[j] + d.cout, m.cin, skip, net.C[l], n, net.C[l], st, join2 ? arena + P.catr[j] + d.cout : nullptr,
                          m.cin));
        const Out2 tail = (dual && j + 1 < nd) ? Out2{arena + P.rl[j + 1], RR} : NO_OUT2;
        bool r_ready = false;
        if (!tf32) {
            SCN_TRY(scn_conv_layer_fwd(arena + P.cat[j], m.cin, n, m.cin, 0, arena + P.catr[j], nullptr, n, 1, m.w, m.img_f, 0, m.b,
                                       arena + P.nin[j], m.cout, tf32, stream));
        } else if (n > 0) {
            if (!join2) SCN_TRY(scn_round_tf32(arena + P.cat[j], arena + P.catr[j], (int64_t)n * m.cin, stream));
            const Out2 o = !dual ? NO_OUT2 : (U ? Out2{arena + P.dr[j][0], RR} : tail);
            SCN_TRY(scn_conv_fwd_tf32_dual(arena + P.catr[j], m.cin, m.cin, n, nullptr, n, 1, m.img_f, m.b, nullptr, 0, nullptr, 0,
                                           arena + P.nin[j], m.cout, m.cout, 0, o.p, m.cout, o.epi, stream));
            r_ready = dual && U > 0;
        }
        SCN_TRY(stage_fwd(net.du[j], U, arena + P.nin[j], r_ready, n, net.CD[j], g.subm[l], arena, P.dr[j], P.dh[j], P.dy[j], dual, tail,
                          tf32, stream));
        cur = arena + P.D[j];
        operand_ready = dual && tail.p && n > 0;
    }
    return SCN_OK;
}

// seeds[k] (k as in scn_unet_plan's output order): incoming gradient of output k or NULL.  pgrads: one pointer per parameter in
// table order (encoder level: entry w, b, then per unit w1 b1 w2 b2; decoder level: deconvolution w, b, 1x1 w, b, units), NULL =
// not wanted; every gradient is ADDED to its buffer (zeroed by the caller, or the parameter's gradient bucket).  gx: gradient
// wrt the network input or NULL.  phases: bit 0 = decoder, bit 1 = encoder levels >= split, bit 2 = encoder levels < split,
// split = phases >> 8 (several calls let the caller start the allreduce of a finished group's gradients while the rest of the
// backward runs; every call must see the same seeds).
int scn_unet_bwd(const int64_t* net_table, const int64_t* geo_table, const float* x, const float* arena, float* barena,
                 const int64_t* seeds, const int64_t* pgrads, float* gx, int phases, int use_tf32, scn_stream_t stream) {
    SCN_REQUIRE(net_table && geo_table && arena && barena && seeds && pgrads, "unet_bwd: null argument");
    Net net;
    SCN_TRY(parse_net(net_table, net));
    Geo g;
    parse_geo(geo_table, net, g);
    Plan P;
    make_plan(net, g, P);
    cudaStream_t st = as_stream(stream);
    const int tf32 = use_tf32 ? 1 : 0, L = net.L;
    // phases: bit 0 = decoder, bit 1 = encoder levels >= split, bit 2 = encoder levels < split, split = phases >> 8
    const int split = phases >> 8;
    const bool run_dec = phases & 1;
    auto enc_runs = [&](int level) { return (phases & (level >= split ? 2 : 4)) != 0; };
    Side side;
    SCN_TRY(side_for(st, side));
    // programmatic dependent launch per stream while two streams feed the GPU (common.cuh: PdlMask).  SCN_EXEC_BWD_PDL:
    // bit 0 = main stream keeps the attribute, bit 1 = side stream keeps it
    PdlMaskScope pdl_scope;
    if (side.on) {
        const char* e = getenv("SCN_EXEC_BWD_PDL");
        const int keep = e ? atoi(e) : SCN_EXEC_BWD_PDL_DEFAULT;
        if (!(keep & 1)) pdl_scope.exclude(side.main);
        if (!(keep & 2))
            for (int i = 0; i < side.n_side; ++i) pdl_scope.exclude(side.sides[i]);
    }
    // parameter-gradient pointers in table order
    float* pg_enc[MAX_LEVELS][2 + 4 * MAX_UNITS];
    float* pg_dec[MAX_LEVELS][4 + 4 * MAX_UNITS];
    {
        const int64_t* q = pgrads;
        for (int i = 0; i < L; ++i) {
            if (net.enc[i].kind) pg_enc[i][0] = as_ptr<float>(*q++), pg_enc[i][1] = as_ptr<float>(*q++);
            else pg_enc[i][0] = pg_enc[i][1] = nullptr;
            for (int u = 0; u < 4 * net.enc_units[i]; ++u) pg_enc[i][2 + u] = as_ptr<float>(*q++);
        }
        for (int j = 0; j < L - 1; ++j)
            for (int u = 0; u < 4 + 4 * net.dec_units[j]; ++u) pg_dec[j][u] = as_ptr<float>(*q++);
    }
    // gradient of an output: nothing yet / the caller's seed (read only) / this call's buffer
    struct Grad {
        const float* seed;
        float* buf;
        bool have;      // buf holds a value
        bool exact;     // ... that is TF32-representable (single contribution written by the rounding ReLU backward)
        bool rounded;   // round(buf) already sits in the last unit's gyr scratch (second output of the fused add below)
    };
    Grad gE[MAX_LEVELS], gD[MAX_LEVELS];
    for (int i = 0; i < L; ++i) gE[i] = Grad{as_ptr<const float>(seeds[i]), barena + P.gE[i], false, false, false};
    for (int j = 0; j < L - 1; ++j) gD[j] = Grad{as_ptr<const float>(seeds[L + j]), barena + P.gD[j], false, false, false};
    // add a contribution that a kernel is about to write: returns where to write it; `commit` folds it in afterwards
    auto target = [&](Grad& G, float* tmp) -> float* { return G.have ? tmp : G.buf; };
    auto commit = [&](Grad& G, float* written, int64_t count, bool run, bool written_exact = false) -> int {
        if (G.have) {
            if (run && count > 0) SCN_TRY(scn_add(G.buf, written, G.buf, count, stream));
            G.exact = false;
        } else {
            G.have = true, G.exact = written_exact;
        }
        return SCN_OK;
    };
    // the value to back-propagate: seed + buffer
    auto resolve = [&](Grad& G, int64_t count, bool run, const float** out) -> int {
        if (G.have && G.seed) {
            if (run && count > 0) SCN_TRY(scn_add(G.buf, G.seed, G.buf, count, stream));
            G.exact = false;
            *out = G.buf;
        } else {
            if (!G.have) G.exact = false;      // the caller's seed: any fp32 value
            *out = G.have ? G.buf : G.seed;      // may be NULL: no gradient reaches this output
        }
        return SCN_OK;
    };

    const bool dual = tf32 && dual_outputs();
    for (int j = L - 2; j >= 0; --j) {
        const int l = L - 2 - j, n = g.n[l], n_in = g.n[l + 1], c = net.CD[j];
        const Conv &d = net.deconv[j], &m = net.nin[j];
        const float* gy = nullptr;
        SCN_TRY(resolve(gD[j], (int64_t)n * c, run_dec, &gy));
        if (!gy) continue;
        Grad& below = j == 0 ? gE[L - 1] : gD[j - 1];      // gradient of this level's input
        if (run_dec) {
            const float* g_nin = gy;
            const int U = net.dec_units[j];
            // unit 0 hands round(gradient of the 1x1 layer's output) to that layer; the 1x1 layer hands round(gradient of the
            // joined columns) to the transposed convolution, which reads its left columns in place (no copy, no rounding pass)
            SCN_TRY(stage_bwd(net.du[j], U, gy, n, c, g.subm[l], arena, P.dr[j], P.dh[j], barena, P.d_gyr[j], P.d_gh[j], P.d_gx[j],
                              pg_dec[j] + 4, dual, dual ? Out2{barena + P.d_round_nin[j], SCN_EPI_ROUND} : NO_OUT2, (tf32 && gD[j].exact) ? 2 : 0, tf32,
                              stream, side, &g_nin));
            // 1x1 layer over the joined columns: input gradient [n, cin], weight + bias gradient
            SCN_TRY(conv_bwd(g_nin, n, m.cout, dual && U > 0, barena + P.d_round_nin[j], m.cout, arena + P.cat[j], m.cin, n, m.cin, nullptr,
                             nullptr, 1, m.w, m.img_b, 0, barena + P.g_cat[j], dual ? Out2{barena + P.g_catr[j], SCN_EPI_ROUND} : NO_OUT2,
                             pg_dec[j][2], pg_dec[j][3], tf32, stream, side));
            // split the joined gradient: left columns -> transposed convolution, right columns -> the skip connection
            if (!dual) SCN_TRY(copy_cols(barena + P.g_up[j], d.cout, barena + P.g_cat[j], m.cin, n, d.cout, st));
        }
        {
            float* skip_to = target(gE[l], barena + P.g_skip[j]);
            if (run_dec) SCN_TRY(copy_cols(skip_to, net.C[l], barena + P.g_cat[j] + d.cout, m.cin, n, net.C[l], st));
            SCN_TRY(commit(gE[l], skip_to, (int64_t)n * net.C[l], run_dec));
        }
        if (run_dec) {
            // transposed convolution backward: the input gradient runs over cmap (children of each coarse row)
            if (dual)      // ... with the backward of the ReLU in front of it (mask = its output, rounded result) in the epilogue
                SCN_TRY(conv_bwd(barena + P.g_cat[j], n, d.cout, n > 0, barena + P.g_catr[j], m.cin, arena + P.rl[j], d.cin, n_in, d.cin,
                                 g.dmap[l], g.cmap[l], d.K, d.w, d.img_b, 0, target(below, barena + P.tmp[l + 1]), NO_OUT2, pg_dec[j][0],
                                 pg_dec[j][1], tf32, stream, side, nullptr, arena + P.rl[j]));
            else
                SCN_TRY(conv_bwd(barena + P.g_up[j], n, d.cout, false, barena + P.d_round_up[j], d.cout, arena + P.rl[j], d.cin, n_in, d.cin,
                                 g.dmap[l], g.cmap[l], d.K, d.w, d.img_b, 0, barena + P.g_rl[j], NO_OUT2, pg_dec[j][0], pg_dec[j][1], tf32,
                                 stream, side));
        }
        {
            float* to = target(below, barena + P.tmp[l + 1]);
            if (run_dec && !dual) SCN_TRY(scn_relu_bwd(arena + P.rl[j], barena + P.g_rl[j], to, (int64_t)n_in * d.cin, tf32, stream));
            SCN_TRY(commit(below, to, (int64_t)n_in * d.cin, run_dec, tf32 != 0));      // k_relu_bwd<true> rounds what it writes
        }
    }
    for (int i = L - 1; i >= 0; --i) {
        const Conv& e = net.enc[i];
        const int n = g.n[i], n_in = i == 0 ? g.n[0] : g.n[i - 1];
        const bool run_enc = enc_runs(i);
        const int U = net.enc_units[i];
        const float* gy = nullptr;
        SCN_TRY(resolve(gE[i], (int64_t)n * net.C[i], run_enc, &gy));
        if (!gy) continue;
        const float* g_c = gy;
        if (run_enc)
            SCN_TRY(stage_bwd(net.eu[i], U, gy, n, net.C[i], g.subm[i], arena, P.er[i], P.eh[i], barena, P.e_gyr[i], P.e_gh[i], P.e_gx[i],
                              pg_enc[i] + 2, dual, (dual && e.kind) ? Out2{barena + P.e_round[i], SCN_EPI_ROUND} : NO_OUT2, (tf32 && gE[i].exact) ? 2 : ((tf32 && gE[i].rounded) ? 1 : 0),
                              tf32, stream, side, &g_c));
        if (!e.kind) {      // pass-through level 0: its gradient IS the input gradient
            if (gx && run_enc && n > 0) {
                cudaError_t err = cudaMemcpyAsync(gx, g_c, (size_t)n * net.C[i] * 4, cudaMemcpyDeviceToDevice, st);
                SCN_REQUIRE(err == cudaSuccess, "unet_bwd: copy of the input gradient: %s", cudaGetErrorString(err));
            }
            continue;
        }
        const int32_t* fmap = e.kind == 2 ? g.cmap[i - 1] : (e.K == 1 ? nullptr : g.subm[i]);
        const int32_t* bmap = e.kind == 2 ? g.dmap[i - 1] : fmap;
        // the layer's input as the forward received it (the weight-gradient MMAs truncate it, exactly like the per-layer path
        // that saves the unrounded tensor); xr / catr are only the forward's rounded operands
        const float* x_in = (i == 0 || P.E[i - 1] < 0) ? x : arena + P.E[i - 1];
        const int reverse = e.kind == 1 ? 1 : 0;
        const bool go_exact = dual && U > 0 && n > 0;
        if (i == 0) {
            if (run_enc)
                SCN_TRY(conv_bwd(g_c, n, e.cout, go_exact, barena + P.e_round[i], e.cout, x_in, e.cin, n_in, e.cin, fmap, bmap, e.K, e.w,
                                 e.img_b, reverse, gx, NO_OUT2, pg_enc[i][0], pg_enc[i][1], tf32, stream, side));
        } else if (dual && gE[i - 1].have) {
            // the skip connection's gradient already sits in gE[i-1]: add it in this convolution's epilogue, in place, and
            // hand round(sum) to the last unit of level i-1 as a second output (nothing else adds to it: no seed)
            Grad& G = gE[i - 1];
            const int Ub = net.enc_units[i - 1];
            const bool hand = Ub > 0 && !G.seed && n_in > 0;
            if (run_enc)
                SCN_TRY(conv_bwd(g_c, n, e.cout, go_exact, barena + P.e_round[i], e.cout, x_in, e.cin, n_in, e.cin, fmap, bmap, e.K, e.w,
                                 e.img_b, reverse, G.buf, hand ? Out2{barena + P.e_gyr[i - 1][Ub - 1], SCN_EPI_ROUND} : NO_OUT2, pg_enc[i][0],
                                 pg_enc[i][1], tf32, stream, side, G.buf));
            G.exact = false, G.rounded = hand;
        } else {
            float* to = target(gE[i - 1], barena + P.tmp[i - 1]);
            if (run_enc)
                SCN_TRY(conv_bwd(g_c, n, e.cout, go_exact, barena + P.e_round[i], e.cout, x_in, e.cin, n_in, e.cin, fmap, bmap, e.K, e.w,
                                 e.img_b, reverse, to, NO_OUT2, pg_enc[i][0], pg_enc[i][1], tf32, stream, side));
            SCN_TRY(commit(gE[i - 1], to, (int64_t)n_in * e.cin, run_enc));
        }
    }
    return side.join();      // the caller's stream continues (optimizer, allreduce) only after every weight gradient
}

}  // extern "C"
