/*
 * scn_b200.h -- C ABI of the B200-native sparse-convolution hot path.
 *
 * Drop-in boundary: the reference's ndsis/modules import the Python
 * namespace `sparseconvnet` (module_factory.py:5, model.py:6, custom_operations.py:4,
 * roi_select_sparse.py:3).  Upstream SparseConvNet binds that namespace to native code
 * through a pybind11 module `sparseconvnet.SCN` whose free functions take a `Metadata<d>&`
 * plus at::Tensors (`X_updateOutput / X_updateGradInput / X_backward`).  This header is
 * what that binding layer is replaced with: plain pointers + sizes + a CUDA stream, no
 * torch types.  Every entry point cites the reference call site it serves.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host;
 *   - features are row-major fp32 [rows, C] with an explicit leading dimension `ld` (floats);
 *   - "keys" are packed voxel coordinates  b<<48 | x<<32 | y<<16 | z  (each field 16 bit);
 *   - neighbour maps are offset-major int32 [K, n_out], -1 = inactive ("output-stationary
 *     rulebook": map[o][r] is the input row that contributes to output row r through
 *     kernel offset o; offsets enumerate the filter box with the LAST dimension fastest,
 *     exactly like SparseConvNet's rulebooks);
 *   - nothing here allocates: the caller owns every buffer (the Python host uses the torch
 *     caching allocator); functions only enqueue work on `stream`;
 *   - return value 0 = ok, otherwise an SCN_ERR_* code; scn_last_error() gives the text
 *     (thread-local).  The library never aborts the process.
 */
#ifndef SCN_B200_H_
#define SCN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* scn_stream_t; /* cudaStream_t */

#define SCN_OK 0
#define SCN_ERR_INVALID 1 /* bad argument (shape / alignment / unsupported size) */
#define SCN_ERR_CUDA 2    /* a CUDA runtime call failed (incl. out of memory)   */

#define SCN_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define SCN_SCAN_BLOCK 4096 /* items per scan block; tmp needs scn_scan_tmp_elems(n) ints */

/* epilogue flags for the convolution kernels */
/* order of application: bias, MASK, ADD, RELU, ROUND */
#define SCN_EPI_RELU 1      /* out = max(out, 0)                                        */
#define SCN_EPI_ADD 2       /* out += residual[row] (same shape, leading dim ld_res)    */
#define SCN_EPI_MASK 4      /* out = mask[row][col] > 0 ? out : 0  (ReLU backward)      */
#define SCN_EPI_ROUND 8     /* out = rna_tf32(out): ready to be a tcgen05 TF32 operand  */

const char* scn_last_error(void);
int scn_version(void);
/* number of SMs / tcgen05 availability of the current device (0 if no device). */
int scn_device_sm_count(void);
int scn_device_is_sm100(void);
/* number of kernels this library has launched in this process (monotonic; bench accounting) */
int64_t scn_launch_count(void);

/* ------------------------------------------------------------------ rulebook builder ------
 * Replaces SparseConvNet Metadata<d> (CPU hash maps).  Call sites: CustomInputLayer.forward
 * custom_operations.py:67-86; RawToTensor/TensorToTensor combiners roi_select_sparse.py:75-84,
 * 113-122; get_spatial_locations custom_operations.py:26-30, roi_select_sparse.py:103. */

/* coords int64 [P, ncol] (x,y,z[,b]); ncol 3 => b = 0.  err_flag (device int, caller zeroes it)
 * is set to 1 if a coordinate is outside [0, 65534]. */
int scn_pack_coords(const int64_t* coords, int P, int ncol, uint64_t* keys, int* err_flag,
                    scn_stream_t stream);
/* inverse: keys -> int64 [n,4] (x,y,z,b)  (get_spatial_locations) */
int scn_unpack_keys(const uint64_t* keys, int n, int64_t* coords, scn_stream_t stream);

/* open-addressing table: tab_keys[cap] (uint64), tab_vals[cap] (int32); cap = power of two */
int scn_hash_clear(uint64_t* tab_keys, int32_t* tab_vals, uint32_t cap, scn_stream_t stream);
/* insert keys[i] -> min(i)  (value = index of first appearance) */
int scn_hash_insert_first(const uint64_t* keys, int P, uint64_t* tab_keys, int32_t* tab_vals,
                          uint32_t cap, scn_stream_t stream);
/* first[i] = 1 iff point i is the first appearance of its voxel */
int scn_hash_first_flags(const uint64_t* keys, int P, const uint64_t* tab_keys,
                         const int32_t* tab_vals, uint32_t cap, int32_t* first,
                         scn_stream_t stream);
/* exclusive prefix sum over n int32; out has n+1 entries (out[n] = total). */
int64_t scn_scan_tmp_elems(int64_t n);
int scn_exclusive_scan(const int32_t* in, int32_t* out, int64_t n, int32_t* tmp,
                       scn_stream_t stream);
/* point_row[i] = rank[first point of i's voxel]; row_keys[rank[i]] = keys[i] for first points.
 * rank = exclusive scan of first[]. */
int scn_hash_assign_rows(const uint64_t* keys, int P, const uint64_t* tab_keys,
                         const int32_t* tab_vals, uint32_t cap, const int32_t* rank,
                         int32_t* point_row, uint64_t* row_keys, scn_stream_t stream);
/* tab_vals[s] = rank[tab_vals[s]] for occupied slots: the table now maps key -> row */
int scn_hash_finalize(const uint64_t* tab_keys, int32_t* tab_vals, uint32_t cap,
                      const int32_t* rank, scn_stream_t stream);
/* Spatially coherent row order (round 2, csrc/sort.cu; no reference counterpart: SparseConvNet numbers level-0 rows by first
 * appearance, Metadata/InputLayer.h, and SURVEY 8c admits any batch-sorted order).  Stable radix sort of the points by
 * (b, Morton(x, y, z)): perm[i] = index of the i-th point in sorted order, sorted_keys[i] = keys[perm[i]].  Only digits
 * that can be non-zero are sorted: coord_bits = bits of the largest coordinate (from the spatial size), batch_bits likewise
 * for the sample index.  ws: scn_morton_order_ws_bytes(P) bytes of scratch.  The level builder then runs over
 * sorted_keys (first appearance over the sorted list = Morton order of the voxels) and scn_scatter_i32 returns its
 * point_row to the original point order: dst[perm[i]] = src[i]. */
int64_t scn_morton_order_ws_bytes(int P);
int scn_morton_order(const uint64_t* keys, int P, int coord_bits, int batch_bits, int32_t* perm,
                     uint64_t* sorted_keys, void* ws, scn_stream_t stream);
int scn_scatter_i32(const int32_t* src, const int32_t* perm, int n, int32_t* dst, scn_stream_t stream);
/* One level around its single host round trip, as two calls: scn_level_count = clear + insert_first + first_flags +
 * exclusive scan (rank [P + 1]; rank[P] is the number of active rows the host reads back), scn_level_finish =
 * assign_rows + finalize.  Same kernels as the separate entry points above. */
int scn_level_count(const uint64_t* keys, int P, uint64_t* tab_keys, int32_t* tab_vals, uint32_t cap,
                    int32_t* first, int32_t* rank, int32_t* scan_tmp, scn_stream_t stream);
int scn_level_finish(const uint64_t* keys, int P, uint64_t* tab_keys, int32_t* tab_vals, uint32_t cap,
                     const int32_t* rank, int32_t* point_row, uint64_t* row_keys, scn_stream_t stream);
/* rows[i] = lookup(keys[i]) or -1 */
int scn_hash_lookup(const uint64_t* keys, int n, const uint64_t* tab_keys,
                    const int32_t* tab_vals, uint32_t cap, int32_t* rows, scn_stream_t stream);

/* CSR of the input rule (row -> its points, ascending point index): counts then fill.
 * row_cnt[N] must be zeroed by the caller; row_ptr = exclusive scan of row_cnt (N+1);
 * cursor[N] zeroed by the caller.  SparseConvNet InputLayer.h rule rows [count, idx...]. */
int scn_rule_count(const int32_t* point_row, int P, int32_t* row_cnt, scn_stream_t stream);
int scn_rule_fill(const int32_t* point_row, int P, const int32_t* row_ptr, int32_t* cursor,
                  int32_t* row_pts, scn_stream_t stream);
int scn_rule_sort(const int32_t* row_ptr, int N, int32_t* row_pts, scn_stream_t stream);
/* count + exclusive scan + fill + sort as ONE call (same kernels): row_ptr [N + 1], row_pts [P]; cursor: N int32 and
 * scan_tmp: scn_scan_tmp_elems(N) int32 of scratch. */
int scn_input_rule(const int32_t* point_row, int P, int N, int32_t* cursor, int32_t* row_ptr,
                   int32_t* row_pts, int32_t* scan_tmp, scn_stream_t stream);

/* submanifold neighbour map for an (fx,fy,fz) filter (odd sizes): map [fx*fy*fz, N].
 * SubmanifoldConvolution call sites module_factory.py:377-414. */
int scn_subm_map(const uint64_t* row_keys, int N, const uint64_t* tab_keys,
                 const int32_t* tab_vals, uint32_t cap, int fx, int fy, int fz, int32_t* map,
                 scn_stream_t stream);

/* strided level (filter == stride): parent key and in-box offset index of every fine row.
 * Convolution/Deconvolution/pooling call sites module_factory.py:221-271,315-354. */
int scn_stride_keys(const uint64_t* row_keys, int N, int sx, int sy, int sz,
                    uint64_t* parent_keys, int32_t* offs, scn_stream_t stream);
/* cmap [K, n_out] (caller pre-fills with -1): cmap[offs[i]][parent_row[i]] = i
 * dmap [K, n_in]: dmap[o][i] = (o == offs[i]) ? parent_row[i] : -1 */
int scn_strided_maps(const int32_t* parent_row, const int32_t* offs, int n_in, int n_out, int K,
                     int32_t* cmap, int32_t* dmap, scn_stream_t stream);
/* A strided level as two calls around its host round trip (same kernels as scn_stride_keys, scn_level_count /
 * scn_level_finish and scn_strided_maps; cmap [K, n_out] is cleared to -1 by the second call). */
int scn_strided_level_count(const uint64_t* fine_keys, int n_in, int sx, int sy, int sz,
                            uint64_t* parent_keys, int32_t* offs, uint64_t* tab_keys, int32_t* tab_vals,
                            uint32_t cap, int32_t* first, int32_t* rank, int32_t* scan_tmp,
                            scn_stream_t stream);
int scn_strided_level_finish(const uint64_t* parent_keys, const int32_t* offs, int n_in, uint64_t* tab_keys,
                             int32_t* tab_vals, uint32_t cap, const int32_t* rank, int32_t* parent_row,
                             uint64_t* row_keys, int n_out, int K, int32_t* cmap, int32_t* dmap,
                             scn_stream_t stream);

/* ------------------------------------------------------------------ convolution family -----
 * out[r] = bias + sum_o  in[map[o][r]] . W[o]      (SubmanifoldConvolution, Convolution,
 * Deconvolution, NetworkInNetwork and all their input-gradients share this form).
 * map == NULL means K == 1 with the identity map (NetworkInNetwork, module_factory.py:357-374). */

/* bytes of the packed weight image for the tcgen05 kernel */
int64_t scn_conv_weight_image_bytes(int K, int Cin, int Cout);
/* Cin/Cout are the GEMM widths of the image (reduction width, output width).  transpose=0: w is
 * [K, Cin, Cout] (SparseConvNet layout [K, 1, Cin, Cout]); transpose=1: w is [K, Cout, Cin] and
 * W[o]^T is packed (input gradients); reverse=1 packs offset K-1-o at o (input gradient of a
 * submanifold convolution).  Values are rounded to TF32 (rna). */
int scn_conv_pack_weights(const float* w, int K, int Cin, int Cout, int transpose, int reverse,
                          void* image, scn_stream_t stream);
/* The same for n images in ONE launch (after optimizer.step() every layer needs both orientations re-packed:
 * ~120 launches per training step otherwise).  table: device array of n rows of 7 int64
 * {w pointer, image pointer, K, Cin, Cout, transpose, reverse}. */
int scn_conv_pack_weights_multi(const int64_t* table, int n, scn_stream_t stream);
/* Tile book of a 3x3x3 submanifold neighbour map over spatially coherent (Morton-ordered) rows (round 2, csrc/conv_ts.cu;
 * the reference side is the same SubmanifoldConvolution call sites, module_factory.py:377-414): for every 128-row output
 * tile the list of DISTINCT input rows its 27 x 128 neighbour references touch, the map re-expressed as 16-bit indices into
 * that list, and a bit mask of offsets with at least one active pair.  scn_tile_book_attach associates a built book with
 * the device pointer of its map: scn_conv_fwd_tf32 (and the fused entry points on top of it) then run the tile-local
 * tensor-memory kernel for that map when the layer qualifies (C in {16, 32, 48, 64}, >= 2 tiles per SM) -- same results up
 * to fp32 summation order.  The caller owns both buffers and detaches before freeing them. */
int64_t scn_tile_book_bytes(int n_out);
int scn_tile_book_build(const int32_t* map, int n_out, int K, void* book, scn_stream_t stream);
int scn_tile_book_attach(const int32_t* map, const void* book, int n_out);
int scn_tile_book_detach(const int32_t* map);
/* detach only if `map` is still associated with `book` (a caller that recycles device addresses) */
int scn_tile_book_detach_if(const int32_t* map, const void* book);
/* launches of the tile-local kernel so far (tests assert that the path under test really ran) */
int64_t scn_conv_ts_launch_count(void);
/* launches of the tile-local weight-gradient kernel (csrc/conv_wgrad_ts.cu) so far */
int64_t scn_conv_wgrad_ts_launch_count(void);
/* TF32 tcgen05 implicit gather-GEMM (sm_100a).  Cin/Cout here are the GEMM's K/N widths, i.e.
 * after any transpose; n_in = rows of `in`.  residual (may be NULL) is [n_out, Cout] with leading
 * dimension ld_res.  When `in` and its row stride are 16-byte aligned the rows are gathered by TMA
 * (tile::gather4, fp32->TF32 round-to-nearest on load), otherwise by cp.async (operand truncated:
 * round it first with scn_round_tf32). */
int scn_conv_fwd_tf32(const float* in, int ld_in, int Cin, int n_in, const int32_t* map, int n_out, int K,
                      const void* image, const float* bias, const float* residual, int ld_res,
                      const float* mask, int ld_mask, float* out, int ld_out, int Cout, int epi_flags,
                      scn_stream_t stream);
/* Same with a second output (out2 may be NULL): out2 = epi2(out), epi2_flags a subset of SCN_EPI_RELU | SCN_EPI_ROUND.
 * In the reference's graph (module_factory.py:127-183) every residual unit starts with scn.ReLU on a tensor whose
 * unrectified value the AddTable shortcut still needs; the producing convolution writes both instead of a separate
 * elementwise pass per unit. */
int scn_conv_fwd_tf32_dual(const float* in, int ld_in, int Cin, int n_in, const int32_t* map, int n_out, int K,
                           const void* image, const float* bias, const float* residual, int ld_res,
                           const float* mask, int ld_mask, float* out, int ld_out, int Cout, int epi_flags,
                           float* out2, int ld_out2, int epi2_flags, scn_stream_t stream);
/* exact fp32 FFMA path (verification mode, <=1e-5).  w is the raw [K, Cin_w, Cout_w] tensor;
 * transpose/reverse as above. */
int scn_conv_fwd_fp32(const float* in, int ld_in, int Cin, const int32_t* map, int n_out, int K,
                      const float* w, int transpose, int reverse, const float* bias,
                      const float* residual, int ld_res, const float* mask, int ld_mask, float* out,
                      int ld_out, int Cout, int epi_flags, scn_stream_t stream);
/* grad_w[o] += in[map[o][:]]^T . grad_out   (fp32 accumulate: grad_w must be zeroed or hold earlier contributions).
 * grad_bias (may be NULL): grad_bias[c] += sum_r grad_out[r][c], computed from the grad-out tiles the kernel stages
 * anyway (saves a pass over grad_out and a launch per layer). */
int scn_conv_bwd_weight(const float* in, int ld_in, int Cin, const int32_t* map, int n_out,
                        int K, const float* grad_out, int ld_go, int Cout, float* grad_w,
                        float* grad_bias, int use_tf32, scn_stream_t stream);
/* Residual unit  y = x + conv2(relu(conv1(relu(x))))  (module_factory.py:127-183, relu_first, identity shortcut;
 * both convolutions C -> C over the same map) as ONE call that enqueues all kernels (relu+round, two gather-GEMMs
 * with fused ReLU / residual-add epilogues); r = relu(x) and h = relu(conv1) are kept for the backward.  img1/img2:
 * caller-owned packed-weight buffers (scn_conv_weight_image_bytes), re-packed here when repack != 0.  All tensors
 * are dense [n, C] with leading dimension C. */
int scn_residual_unit_fwd(const float* x, int n, int C, const int32_t* map, int K, const float* w1,
                          const float* b1, const float* w2, const float* b2, void* img1, void* img2,
                          int repack, float* r, float* h, float* y, int use_tf32, scn_stream_t stream);
/* backward of the unit: gx = gy + relu'(x) * conv1^T(relu'(h) * conv2^T(gy)); weight / bias gradients (any of the
 * g* outputs may be NULL).  gyr [n, C] is scratch (TF32-rounded gy), gh [n, C] receives d/dh.  accumulate != 0:
 * gw* / gb* are ADDED to (the parameters' gradient buffers, training.py:458), otherwise they are overwritten. */
int scn_residual_unit_bwd(const float* gy, const float* r, const float* h, int n, int C, const int32_t* map,
                          int K, const float* w1, const float* w2, void* img1t, void* img2t, int repack,
                          float* gyr, float* gh, float* gx, float* gw1, float* gb1, float* gw2, float* gb2,
                          int accumulate, int use_tf32, scn_stream_t stream);
/* One convolution layer (SubmanifoldConvolution / Convolution / Deconvolution / NetworkInNetwork, module_factory.py:221-414)
 * per direction as ONE call that enqueues: TF32 rounding of the gathered operand unless *_exact (scratch x_round / go_round),
 * weight packing when repack != 0, the gather-GEMM(s), and the weight + bias gradient (gw, gb are ADDED to: zeroed buffers
 * or the parameters' gradient buckets).  bwd: fmap [K, n_out] is the forward map, bmap [K, n_in] the map of the input
 * gradient, reverse_bwd = 1 for submanifold layers (bmap == fmap, offsets reversed); any of gx / gw / gb may be NULL. */
int scn_conv_layer_fwd(const float* x, int ld_x, int n_in, int Cin, int x_exact, float* x_round,
                       const int32_t* map, int n_out, int K, const float* w, void* image, int repack,
                       const float* bias, float* out, int Cout, int use_tf32, scn_stream_t stream);
int scn_conv_layer_bwd(const float* go, int n_out, int Cout, int go_exact, float* go_round, const float* x,
                       int ld_x, int n_in, int Cin, const int32_t* fmap, const int32_t* bmap, int K,
                       const float* w, void* image_t, int repack, int reverse_bwd, float* gx, float* gw,
                       float* gb, int use_tf32, scn_stream_t stream);
/* out[c] = sum_r in[r][c]   (bias gradient) */
int scn_col_sum(const float* in, int ld, int n, int C, float* out, scn_stream_t stream);
/* out[c] += sum_r in[r][c]  (bias gradient accumulated straight into the parameter's .grad buffer) */
int scn_col_sum_add(const float* in, int ld, int n, int C, float* out, scn_stream_t stream);

/* ------------------------------------------------------------------ elementwise ------------
 * scn.ReLU module_factory.py:86-89; AddTable :51-57; BatchNorm(Leaky)ReLU :92-113. */
/* round_tf32 != 0: the result is additionally rounded (to nearest, ties away) to TF32 so that a
 * following tcgen05 kind::tf32 MMA, which TRUNCATES fp32 operands, reads it exactly. */
int scn_relu_fwd(const float* in, float* out, int64_t n, int round_tf32, scn_stream_t stream);
int scn_relu_bwd(const float* out_or_in, const float* grad_out, float* grad_in, int64_t n,
                 int round_tf32, scn_stream_t stream);
/* out = rna_tf32(in) */
int scn_round_tf32(const float* in, float* out, int64_t n, scn_stream_t stream);
int scn_add(const float* a, const float* b, float* out, int64_t n, scn_stream_t stream);
/* per-channel mean and biased variance over n rows (two-pass, deterministic) */
int scn_bn_stats(const float* in, int n, int C, float* mean, float* var, scn_stream_t stream);
/* y = leaky((x-mean)*rsqrt(var+eps)*gamma+beta) ; gamma/beta may be NULL */
int scn_bn_apply(const float* in, int n, int C, const float* mean, const float* var,
                 const float* gamma, const float* beta, float eps, float leak, float* out,
                 scn_stream_t stream);
/* backward of train-mode BN+leaky: needs x, y, dy; writes dx and (if non-NULL) dgamma, dbeta */
int scn_bn_bwd(const float* x, const float* y, const float* dy, int n, int C, const float* mean,
               const float* var, const float* gamma, float eps, float leak, int training,
               float* dx, float* dgamma, float* dbeta, float* tmp2C, scn_stream_t stream);

/* ------------------------------------------------------------------ io layers --------------
 * InputLayerFunction custom_operations.py:74-82 (mode 4), roi_select_sparse.py:79-81,117-119;
 * OutputLayerFunction custom_operations.py:7-10, model.py:461,576,600,643,658. */
/* out[r] = reduce over the row's points (mode 1 last, 2 first, 3 sum, 4 mean) */
int scn_input_fwd(const float* feats, int ld, int C, const int32_t* row_ptr,
                  const int32_t* row_pts, int N, int mode, float* out, scn_stream_t stream);
int scn_input_bwd(const float* grad_out, int C, const int32_t* point_row, const int32_t* row_ptr,
                  const int32_t* row_pts, int P, int mode, float* grad_feats,
                  scn_stream_t stream);
/* out[i] = in[idx[i]]  (idx < 0 => zeros)   (OutputLayer forward, crop feature gather) */
int scn_gather_rows(const float* in, int ld_in, const int32_t* idx, int n, int C, float* out,
                    int ld_out, scn_stream_t stream);
/* out[idx[i]] += in[i]  with atomics (backward of gather when no CSR exists) */
int scn_scatter_add_rows(const float* in, const int32_t* idx, int n, int C, float* out,
                         scn_stream_t stream);

/* SparseToDense module_factory.py:429-435: out [B, C, X, Y, Z] incl. the zero fill */
int scn_sparse_to_dense_fwd(const float* in, int C, const uint64_t* tab_keys,
                            const int32_t* tab_vals, uint32_t cap, int B, int X, int Y, int Z,
                            float* out, scn_stream_t stream);
int scn_sparse_to_dense_bwd(const float* grad_dense, const uint64_t* row_keys, int N, int C,
                            int X, int Y, int Z, float* grad_in, scn_stream_t stream);

/* ------------------------------------------------------------------ point-wise loss ---------
 * nn.CrossEntropyLoss(weight, ignore_index, reduction='mean') on the segmentation logits, ndsis/modules/loss.py:95-97:
 * loss = sum_i w[y_i] * (lse_i - x_i[y_i]) / sum_i w[y_i] over rows with y_i != ignore_index.  fwd writes the per-row
 * log-sum-exp (kept for the backward) and stats = {sum of weighted losses, sum of weights} (deterministic reduction);
 * bwd writes dlogits = (*grad_loss) * w[y] / stats[1] * (softmax - onehot), zero for ignored rows.  weight may be NULL. */
int scn_cross_entropy_fwd(const float* logits, int ld, int64_t n, int C, const int64_t* labels, const float* weight,
                          int64_t ignore_index, float* lse, float* row_scratch, float* stats, scn_stream_t stream);
int scn_cross_entropy_bwd(const float* logits, int ld, int64_t n, int C, const int64_t* labels, const float* weight,
                          int64_t ignore_index, const float* lse, const float* stats, const float* grad_loss,
                          float* dlogits, scn_stream_t stream);

/* ------------------------------------------------------------------ per-box mask loss ---
 * MaskLoss._single_sample_loss + the per-box loop of MaskLoss.forward, ndsis/modules/loss.py:284-300:
 * mean_out[s] = mean over the segment's elements of binary_cross_entropy_with_logits(x_i, t_i)   (NaN for an empty
 * segment, as torch's mean of an empty tensor); segments are the boxes' (box, point) rows, seg_ptr [n_seg + 1].
 * bwd: grad_logits_i = grad_mean[s] / count_s * (sigmoid(x_i) - t_i). */
int scn_segment_bce_fwd(const float* logits, const uint8_t* targets, const int32_t* seg_ptr, int n_seg,
                        float* mean_out, scn_stream_t stream);
int scn_segment_bce_bwd(const float* logits, const uint8_t* targets, const int32_t* seg_ptr, int n_seg,
                        const float* grad_mean, float* grad_logits, scn_stream_t stream);

/* ------------------------------------------------------------------ proposal selection ---
 * ProposalSelector.forward ndsis/modules/proposal_selector.py:60-89 + non_maximum_supression ndsis/utils/bbox.py:713-759
 * (IoU as bbox_overlap_unsqueezed_area_start_end :235-240).  boxes: [B, n, 2, 3] fp32 (start, stop), every sample sorted
 * by DESCENDING score; n <= 4096.  keep [B, n]: 1 where the reference's is_maximum is True; keep_idx [B, max_keep]: the
 * first min(kept, max_keep) survivors in score order, counts [B] their number.  workspace: scn_nms3d_workspace_bytes. */
int64_t scn_nms3d_workspace_bytes(int B, int n);
int scn_nms3d(const float* boxes, int B, int n, float thresh, int max_keep, void* workspace, uint8_t* keep,
              int32_t* keep_idx, int32_t* counts, scn_stream_t stream);

/* ------------------------------------------------------------------ pooling ----------------
 * MaxPooling / AveragePooling module_factory.py:315-354 (cmap from scn_strided_maps);
 * SparseGlobalPool custom_operations.py:42-59 (segment mean over batch-sorted rows). */
int scn_pool_fwd(const float* in, int C, const int32_t* cmap, int n_out, int K, int is_max,
                 float inv_volume, float* out, scn_stream_t stream);
int scn_pool_bwd(const float* in, const float* out, const float* grad_out, int C,
                 const int32_t* parent_row, int n_in, int is_max, float inv_volume,
                 float* grad_in, scn_stream_t stream);
int scn_segment_mean_fwd(const float* in, int C, const int32_t* seg_ptr, int n_seg, float* out,
                         scn_stream_t stream);
int scn_segment_mean_bwd(const float* grad_out, int C, const int32_t* seg_ptr, int n_seg,
                         float* grad_in, scn_stream_t stream);
/* seg_ptr[b] = first row whose batch field >= b (rows batch-sorted); n_seg+1 entries */
int scn_batch_offsets(const uint64_t* row_keys, int N, int n_seg, int32_t* seg_ptr,
                      scn_stream_t stream);

/* ------------------------------------------------------------------ mask crop --------------
 * roi_cut / get_inside_indicator / select_features / select_coords roi_select_sparse.py:125-180.
 * boxes int32 [BB, 2, 3] (start, stop; half-open), box_sample[BB]; points given as keys with
 * sample_ptr[B+1] (points are grouped by sample).  Output is in (box, point) order. */
#define SCN_CROP_CHUNK 2048
/* counts [BB, n_chunks] with n_chunks = ceil(max_sample_len / SCN_CROP_CHUNK) */
int scn_crop_count(const uint64_t* keys, const int32_t* sample_ptr, const int32_t* boxes,
                   const int32_t* box_sample, int BB, int n_chunks, int32_t* counts,
                   scn_stream_t stream);
/* offsets = exclusive scan of counts.  Writes sel_pt[total] (source point index), new_keys[total]
 * (xyz absolute, batch field := box id) and, if non-NULL, is_inside [BB, P] bytes. */
int scn_crop_select(const uint64_t* keys, const int32_t* sample_ptr, const int32_t* boxes,
                    const int32_t* box_sample, int BB, int n_chunks, const int32_t* offsets,
                    int P, int32_t* sel_pt, uint64_t* new_keys, uint8_t* is_inside,
                    scn_stream_t stream);

/* ------------------------------------------------------------------ dense stage -------------
 * get_dilation_network module_factory.py:581-611 (SparseToDense + num_dilations x [Conv3d 3^3 'same', dilated + ReLU]),
 * the trunk of the region-proposal network (anchor_network.py:127-219 consumes its [B, C, X, Y, Z] output).  The dense
 * grid is kept channels-last ([B*X*Y*Z, C] rows in (b, x, y, z) order) so that its convolutions run on the gather-GEMM
 * entry points above (scn_conv_fwd_tf32 / _fp32, scn_conv_bwd_weight) over the trivial map built here. */
/* map [27][B*X*Y*Z]: map[o][r] = r + delta(o) * dilation, -1 outside the grid; offsets last-dimension-fastest */
int scn_dense_map(int B, int X, int Y, int Z, int dilation, int32_t* map, scn_stream_t stream);
/* sparse rows -> dense rows, zero fill fused (scn.SparseToDense, module_factory.py:429-435, channels-last) */
int scn_sparse_to_dense_rows_fwd(const float* in, int C, const uint64_t* tab_keys, const int32_t* tab_vals,
                                 uint32_t cap, int B, int X, int Y, int Z, float* out, scn_stream_t stream);
int scn_sparse_to_dense_rows_bwd(const float* grad_dense, const uint64_t* row_keys, int N, int C, int X, int Y,
                                 int Z, float* grad_in, scn_stream_t stream);
/* out[b][c][r] = in[b][r][c]: dense rows <-> [B, C, X*Y*Z] */
int scn_transpose_batched(const float* in, int batches, int rows, int cols, float* out, scn_stream_t stream);

/* ------------------------------------------------------------------ voxelisation + collation ----
 * the deterministic core of the reference's per-sample conversion for a whole batch on the device --
 * augment_coords sparse_augmentation.py:81-126 (coords @ (distortion * scale), shift = -min + sub-pixel offset, .long(),
 * fix_cut_out :41-46 / a drawn random_cut_out :49-78) and collate_fn data.py:88-115 (sample-index column, concatenation).
 * points fp32 [P, 3], samples concatenated, sample_ptr int32 [B + 1] (device); proj fp32 [B, 9] row-major, offset fp32
 * [B, 3]; window int32 [B, 9] = start[3] (inside test: start <= c < start + size), size[3], move[3] (added to the
 * coordinates that stay): fix_cut_out(shift) = {0, size, +shift}, a drawn cut-out = {start, size, -start}.
 * Outputs: coords int64 [P', 4] (x, y, z, sample; buffer of P rows), kept int32 [P'] = input row of every output row,
 * out_ptr int32 [B + 1] = first output row of every sample (out_ptr[B] = P'), shift_out fp32 [B, 3] = -min + offset.
 * fp32 arithmetic is the reference's bit for bit (one rounded product + two fused multiply-adds per output). */
int64_t scn_voxelize_ws_bytes(int P, int B);
int scn_voxelize(const float* points, int P, const int32_t* sample_ptr, int B, const float* proj, const float* offset,
                 const int32_t* window, void* ws, int64_t* coords, int32_t* kept, int32_t* out_ptr, float* shift_out,
                 scn_stream_t stream);
/* augment_features sparse_augmentation.py:129-190 for the kept rows: out [n, C] = [colors (+ color_shift[sample]) | ones |
 * normals @ rotation[sample] (+ normal_shift[sample])]; colors / normals fp32 [P, 3] or NULL, shifts fp32 [B, 3] or NULL,
 * rotation fp32 [B, 9] or NULL; C = 3 * (colors != NULL) + (use_ones != 0) + 3 * (normals != NULL). */
int scn_voxelize_features(const int32_t* kept, int n, const int32_t* out_ptr, int B, const float* colors,
                          const float* color_shift, int use_ones, const float* normals, const float* rotation,
                          const float* normal_shift, float* out, int C, scn_stream_t stream);

/* ------------------------------------------------------------------ sparse U-Net executor ----
 * FeatureExtractor.forward model.py:414-446 over the graph module_factory.py:438-578,789-830 builds (encoder levels:
 * entry convolution + residual units; decoder levels: ReLU, Deconvolution, JoinTable with the skip connection,
 * NetworkInNetwork, residual units) as ONE call per direction that enqueues every kernel of the pass
 * (sparse_rcnn_b200/csrc/unet_exec.cu; host mirror: sparse_rcnn_b200/executor.py).
 *   net : int64 layer table   [L] ; per encoder level [kind K cin cout w b img_f img_b n_units] + n_units x
 *         [w1 b1 w2 b2 img1_f img2_f img1_b img2_b] ; per decoder level [K cin cout w b img_f img_b] (Deconvolution)
 *         [cin cout w b img_f img_b] (NetworkInNetwork) [n_units] + units.  kind: 0 pass-through (level 0 only),
 *         1 submanifold (K = 1 or 27), 2 Convolution 2^3/s2 (K = 8).  img_*: packed weight images (forward / transposed),
 *         packed by the caller (scn_conv_pack_weights_multi); unused in fp32 mode.
 *   geo : int64 geometry table, per level [rows, subm_map(3^3) ptr, cmap ptr (to the next coarser level), dmap ptr].
 * scn_unet_plan: out[0..L) = offsets (floats) of the encoder outputs E_i in the forward arena (-1: the input itself),
 * out[L..2L-1) = decoder outputs D_j, out[2L-1] = forward arena floats, out[2L] = backward arena floats. */
int scn_unet_plan(const int64_t* net, const int64_t* geo, int64_t* out);
int scn_unet_fwd(const int64_t* net, const int64_t* geo, const float* x, float* arena, int n_decoder_levels,
                 int use_tf32, scn_stream_t stream);
/* seeds[k]: incoming gradient of output k (order of scn_unet_plan) or 0.  pgrads: one pointer per parameter in table
 * order (0 = not wanted); gradients are ADDED (zeroed buffers or the parameters' gradient buckets, training.py:458).
 * gx: gradient wrt the input or NULL.  phases: bit 0 decoder, bit 1 encoder levels >= split, bit 2 encoder levels < split,
 * split = phases >> 8 (7 = everything in one call). */
int scn_unet_bwd(const int64_t* net, const int64_t* geo, const float* x, const float* arena, float* bwd_arena,
                 const int64_t* seeds, const int64_t* pgrads, float* gx, int phases, int use_tf32,
                 scn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SCN_B200_H_ */
