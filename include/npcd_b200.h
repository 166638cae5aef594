/* libnpcd_b200 -- C ABI of the B200-native PointNeRF render path (drop-in for the renderer call of
 * lmb-freiburg/neural-point-cloud-diffusion, `npcd/models/pointnerf/pointnerf.py:89-97,126`).
 *
 * Conventions:
 *   - every function returns 0 on success, non-zero on argument (1) or CUDA (2) error; `npcd_last_error()` gives the message.
 *     The reference's convention is Python asserts/exceptions (e.g. `fields/aggregators/aggregator.py:17`,
 *     `renderers/renderer.py:163`); the Python shim turns a non-zero code into `RuntimeError`.
 *   - all buffers are caller-allocated DEVICE pointers (torch tensors' `data_ptr()`), row-major contiguous, fp32 unless noted;
 *     kernels never allocate; `stream` is a `cudaStream_t` (pass `torch.cuda.current_stream().cuda_stream`); calls are
 *     asynchronous with respect to the host.
 *   - no torch types, no global state (apart from a thread-local error string).
 *
 * Reference paths below are relative to /root/reference/npcd/models/pointnerf/.
 */
#ifndef NPCD_B200_H
#define NPCD_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPCD_B200_ABI_VERSION 3

const char* npcd_last_error(void);
int npcd_abi_version(void);

/* ---- R1 + R2: rays and ray/cube limits ------------------------------------------------------------------------------------
 * Replaces RaySampler.forward (renderers/ray_sampler.py:10-63), get_ray_limits_box (renderers/math_utils.py:46-97) and
 * Renderer.get_ray_limits (renderers/renderer.py:36-47), including the train-mode ray subset gather (renderer.py:232-238).
 *   extr [n_views,4,4] world->cam, intr [n_views,3,3]; ray_subset: NULL, or [n_subset] int64 pixel ids shared by all views.
 *   cam_centers [n_views,3]; origins [n_views,R,3] (optional, may be NULL: every ray of a view starts at its camera centre);
 *   dirs [n_views,R,3]; ray_start/ray_end [n_views,R]; limits_scratch: 8 bytes of device scratch.  R = n_subset or res^2.  */
int npcd_rays_generate(const float* extr, const float* intr, int n_views, int resolution, const long long* ray_subset,
                       int n_subset, float cube_scale, float* cam_centers, float* origins, float* dirs, float* ray_start,
                       float* ray_end, void* limits_scratch, void* stream);

/* ---- grid build: replaces torch_knnquery.VoxelGrid.set_pointset (call sites pointnerf.py:67-75,116-124) ---------------------
 *   kp_pos [n_obj,P,3] -> cell_start [n_obj, cells+1] int32, sorted_pts [n_obj,P,4] (x,y,z,index bits),
 *   occ_bits [n_obj, words] uint32; `npcd_grid_dims` reports cells / words.                                                   */
int npcd_grid_dims(int* cells, int* words);
int npcd_grid_build(const float* kp_pos, int n_obj, int n_points, int* cell_start, float* sorted_pts, unsigned* occ_bits,
                    float* aabb /* optional [n_obj,6]: box of the dilated occupied cells, +-inf where it meets the cube border */,
                    void* stream);
/* Sub-cell masks for the marcher (optional): masks [n_obj, cells, 2] uint64 = (sure, maybe) per grid cell, one bit per 1/48-edge
 * sub-cell: a depth sample in a `sure` sub-cell has a point within `radius` for certain, one outside every `maybe` sub-cell has
 * none; only the shell in between takes the exact test, so results are unchanged.  Must be built with the radius later queried.  */
int npcd_grid_build_masks(const float* kp_pos, int n_obj, int n_points, float radius, void* masks, void* stream);

/* ---- march + exact radius-kNN: replaces Aggregator.query_keypoints (fields/aggregators/aggregator.py:25-76) and
 * torch_knnquery.VoxelGrid.query (call site aggregator.py:63), with the depth sampling that feeds them
 * (renderers/renderer.py:49-77, renderers/volume_renderer.py:63-70).
 *   pass 1  npcd_march_count: per ray a 128-bit mask of samples having >= 1 point within `radius` and
 *           ray_count = min(popcount, max_shading_pts).           jitter: NULL or [n_rays,128] U[0,1) (renderer.py:74-76).
 *   scan    npcd_scan_counts: ray_offset[0..n] (int64) over ray_count[ray_ids[i]] (ray_ids NULL = identity).
 *   pass 2  npcd_knn_fill: for the kept samples s < min(ray_offset[n_sel], capacity):
 *           nbr_idx [S,8] int32 (global b*P+p, ascending (distance, index), -1 padded), sample_pos [S,4] = (x,y,z,slot depth),
 *           sample_ray [S] int32 (optional).                                                                                  */
int npcd_march_count(const float* cam_centers, const float* dirs, const float* ray_start, const float* ray_end,
                     const float* jitter, long long n_rays, int rays_per_view, int views_per_obj, int n_points,
                     const int* cell_start, const float* sorted_pts, const unsigned* occ_bits,
                     const float* aabb /* optional, from npcd_grid_build: samples outside the box are skipped untested */,
                     const void* fine_masks /* optional, from npcd_grid_build_masks (same radius) */,
                     float radius, int max_shading_pts, unsigned* valid_bits, int* ray_count,
                     int impl /* 0 auto (n_points <= 2048: march = 2, kNN fill = 3); 1 generic global-memory kernels; 2 shared-memory
                               thread-per-sample kernels (n_points <= 2048); 3 ray-coherent kernels (march: warp per ray with a
                               per-ray candidate list; fill: one candidate list per warp iteration; same results) */,
                     void* stream);
int npcd_scan_workspace_bytes(long long n, size_t* bytes);
int npcd_scan_counts(const int* ray_count, const int* ray_ids, long long n, long long* ray_offset, void* workspace,
                     size_t workspace_bytes, void* stream);
/* Generic exact query on explicit positions x [n,3] (object of query i = query_obj ? query_obj[i] : i / queries_per_obj);
 * used by the TV-loss self-query (npcd/losses/neural_point_cloud_tv_loss.py:41-44) and by Aggregator.query_keypoints. */
int npcd_knn_points(const float* x, const int* query_obj, long long n, int queries_per_obj, int n_points, const int* cell_start,
                    const float* sorted_pts, float radius, int* nbr_idx, void* stream);
int npcd_knn_fill(const float* cam_centers, const float* dirs, const float* ray_start, const float* ray_end, const float* jitter,
                  const int* ray_ids, long long n_sel, const long long* ray_offset, const unsigned* valid_bits, int rays_per_view,
                  int views_per_obj, int n_points, const int* cell_start, const float* sorted_pts, float radius,
                  long long capacity, int* nbr_idx, float* sample_pos,
                  float* sample_t /* optional [capacity]: the slot depths (= sample_pos[.][3]) as a dense array for the compositor */,
                  int* sample_ray, int impl /* as npcd_march_count */, void* stream);

/* ---- Q1: voxel-grid-compatible query mode (the semantics of the reference's deployed path through torch_knnquery.VoxelGrid,
 * fields/aggregators/aggregator.py:59-76 with the options of pointnerf.py:147-153; source un-vendored => PARITY UNPINNED, checked
 * against oracle/pointnerf_oracle.py::query_keypoints_voxel).  Call order: npcd_voxel_select -> npcd_grid_build(stored_pos) ->
 * npcd_march_count(max_shading_pts = 128) -> npcd_voxel_filter -> npcd_scan_counts -> npcd_knn_fill -> npcd_voxel_slots ->
 * field -> npcd_composite_fwd(slot).
 *   npcd_voxel_dims:   voxels per axis (round((hi - lo) / voxel_size), at most 32) and 32-bit words of one object's bit set.
 *   npcd_voxel_select: stored_pos [B,P,3] = the points a voxel keeps (<= max_points_per_voxel per voxel, LOWEST index wins; points
 *                      outside the ranges are dropped too), dropped points moved to a far sentinel npcd_grid_build ignores;
 *                      vox_bits [B,words] = kernel_size^3 dilation of the occupied voxels, bit (x * n + y) * n + z.
 *   npcd_voxel_filter: cand_bits [n_rays,4] = the first max_shading_pts samples of each ray lying in a set voxel; valid_bits &= them;
 *                      ray_count [n_rays] = kept samples (candidates with a stored neighbour within r).
 *   npcd_voxel_slots:  slot [S] u8 = index of every kept sample among its ray's candidates (gaps = holes, aggregator.py:66-70).    */
int npcd_voxel_dims(float voxel_size, float range_lo, float range_hi, int* n_vox, int* words);
int npcd_voxel_select(const float* kp_pos, int n_obj, int n_points, float voxel_size, float range_lo, int n_vox,
                      int max_points_per_voxel, int kernel_size, float* stored_pos, unsigned* vox_bits, void* stream);
int npcd_voxel_filter(const float* cam_centers, const float* dirs, const float* ray_start, const float* ray_end, const float* jitter,
                      long long n_rays, int rays_per_view, int views_per_obj, const unsigned* vox_bits, int n_vox, float voxel_size,
                      float range_lo, int max_shading_pts, unsigned* valid_bits, unsigned* cand_bits, int* ray_count, void* stream);
int npcd_voxel_slots(const long long* ray_offset, const int* ray_ids, const unsigned* valid_bits, const unsigned* cand_bits,
                     long long n_sel, long long capacity, unsigned char* slot, void* stream);

/* ---- Q3: train-mode valid-ray subsampling, replaces Aggregator.subsample_valid_rays (fields/aggregators/aggregator.py:78-119) ------
 *   npcd_count_valid_rays: n_valid [n_views] = #rays of the view with ray_count > 0, min_valid [1] = their minimum (the host reads it:
 *     n = min(min_valid, ray_subsamples) sizes every output, aggregator.py:102-103).
 *   npcd_subsample_valid_rays: per view a uniform random n_keep-subset of its valid rays (counter-based generator keyed by `seed`,
 *     the GLOBAL view number view_offset + view and the draw, so an object-sharded batch picks the same rays as the whole batch
 *     would), written as ray ids view * rays_per_view + ray in ascending order: ray_ids [n_views, n_keep] int32
 *     (n_keep <= every n_valid).                                                                                                   */
int npcd_count_valid_rays(const int* ray_count, long long n_views, int rays_per_view, int* n_valid, int* min_valid, void* stream);
int npcd_subsample_valid_rays(const int* ray_count, long long n_views, int rays_per_view, int n_keep, unsigned long long seed,
                              long long view_offset, int* ray_ids, void* stream);

/* ---- field: gather + posenc + pair MLP + aggregation + density/colour heads -------------------------------------------------
 * Replaces aggregators.MLP.get_local_feat / aggregate_local_feat (fields/aggregators/mlp.py:36-125), Aggregator.get_keypoint_data
 * (aggregator.py:121-144), fields.MLP.get_shape / get_channels (fields/mlp.py:38-72) and the activations of Field.forward
 * (fields/field.py:126-141).  rgbs [S,4] = (r, g, b, sigma).  n_samples_dev: device int64 (= ray_offset + n_sel).
 * fp32 SIMT version: weights pre-transposed to [in, 256] ("wt"), first layer zero-padded to a multiple of 16 rows.            */
typedef struct {
  int feat_dim;
  const float* pair_wt[4]; /* local_field.0,2,4,6 */
  const float* pair_b[4];
  const float* agg_wt; /* local_field.8, applied after the weighted aggregation (weights sum to 1) */
  const float* agg_b;
  const float* shape_wt; /* shape_net.0 */
  const float* shape_b;
  const float* shape_out_w; /* shape_net.2 weight [256] */
  const float* shape_out_b; /* [1] */
  const float* chan_wt[4];  /* channel_net.0,2,4,6 */
  const float* chan_b[4];
  const float* chan_out_w; /* channel_net.8 weight [3,256] */
  const float* chan_out_b; /* [3] */
} npcd_mlp_simt_weights;

int npcd_field_simt_fwd(const int* nbr_idx, const float* sample_pos, const float* kp_pos, const float* kp_feat,
                        const long long* n_samples_dev, long long capacity, const npcd_mlp_simt_weights* weights,
                        float* agg_workspace /* [capacity,256] */, float* rgbs, float* feat_out /* optional [S,256] */,
                        int stages /* bit0: pair MLP + aggregation -> agg_workspace, bit1: heads -> rgbs; 3 = both */,
                        int num_sms, void* stream);

/* Tensor-core (tcgen05 / TMEM) version of the same field.  Each Linear layer is packed ONCE by npcd_tc_pack_weights into
 * pre-swizzled fp16 hi/lo tiles: per 64-column K-block a 32 KB "hi" image then a 32 KB "lo" image of the [256 x 64] weight slice
 * in the K-major SWIZZLE_128B shared-memory layout, multiplied by `scale` (a power of two; pass inv_scale = 1/scale).
 * perm (optional, [k_pad] int32): source column of packed column k (-1 = zero); k_pad: multiple of 16, <= 256.
 * Pair layer 0 uses OUR 112-column input order: [feat 0..31 | x: d, sin f0..9, cos f0..9, 0,0,0 | y: ... | z: ... | 8 zeros].
 * Biases and the two narrow output layers are HOST pointers: they travel to the kernel by value (constant bank).           */
typedef struct {
  const void* packed_w; /* DEVICE: 2 * ceil(k_pad / 64) tiles of 32 KB */
  const float* bias;    /* HOST: [256] */
  float inv_scale;
  int k_pad;
} npcd_tc_layer;

typedef struct {
  int feat_dim;          /* must be 32 */
  npcd_tc_layer pair[4]; /* local_field.0,2,4,6 */
  npcd_tc_layer agg;     /* local_field.8 (after aggregation) */
  npcd_tc_layer shape;   /* shape_net.0 */
  npcd_tc_layer chan[4]; /* channel_net.0,2,4,6 */
  const float* shape_out_w; /* HOST: shape_net.2 weight [256] */
  const float* shape_out_b; /* HOST: [1] */
  const float* chan_out_w;  /* HOST: channel_net.8 weight [3,256] */
  const float* chan_out_b;  /* HOST: [3] */
} npcd_mlp_tc_weights;

int npcd_tc_pack_weights(const float* w /* [256,k_in] */, int k_in, const int* perm, int k_pad, float scale, void* out,
                         void* stream);
/* All weight matrices of one step in one launch (<= 24 jobs): packed element (n, k) = scale * (transpose ? w[k*ld + n] : w[n*ld + k])
 * for n < n_rows and k < k_in (through perm when given), zero elsewhere; same output image as npcd_tc_pack_weights.              */
typedef struct {
  const float* w;
  long long ld;
  int n_rows, k_in, k_pad, transpose;
  const int* perm;
  float scale;
  void* out;
  int format; /* 0: fp16 hi + fp16 lo tiles ("f16x3" scheme); 1: fp16 + e4m3 tiles ("f16+e4m3x2" scheme, see npcd_tc_pack_weights_f8);
                 2: as 1 with the 8-bit tile in K = 32 steps of [Whi8 x 16 | Wlo8 x 16] (pair layers 1..3, npcd_field_tc_fwd stages bit 5) */
} npcd_tc_pack_job;
int npcd_tc_pack_weights_batched(const npcd_tc_pack_job* jobs, int n_jobs, void* stream);
/* workspace (device bytes) for `capacity` kept samples (< 2^27 per launch): the pre-split [S,256] aggregate image that links the
 * pair stage to the heads stage, the dense pair packing (pair offsets, tile starts) and scan scratch.                          */
int npcd_field_tc_workspace_bytes(long long capacity, size_t* bytes);
int npcd_field_tc_fwd(const int* nbr_idx, const float* sample_pos, const float* kp_pos, const float* kp_feat,
                      const long long* n_samples_dev, long long capacity, const npcd_mlp_tc_weights* weights, void* workspace,
                      size_t workspace_bytes, float* rgbs, float* feat_out,
                      int stages /* bit0: dense packing + pair MLP + aggregation -> workspace, bit1: heads -> rgbs, bit2: heads with
                                    local_field.8 folded into W->shape / W->chan[0] by the caller (W' = W W_8, b' = W b_8 + b; W->agg unused),
                                    bit3: `weights` were packed in format 1 -- run the "f16 + e4m3 x 2" operand scheme (one fp16
                                    product + two e4m3 correction products per layer instead of three fp16 products),
                                    bit4 (with bit3): only the activation-rounding correction product is issued ("f16 + e4m3": the
                                    weights are then effectively rounded to fp16; 1.5 instead of 2 tensor passes per product),
                                    bit5 (with bit3, without bit4; pair stage, and heads stage with bit2): weights->pair[1..3] resp.
                                    weights->chan[1..3] were packed in format 2 -- the A operand of those layers lives in tensor
                                    memory (epilogues convert the accumulator in place), shared memory only feeds the weights */,
                      int* error_flag /* device int, optional */, int num_sms, void* stream);
/* fp32 rows [n,256] <-> the pre-split operand image the tensor-core kernels exchange: per 128-row tile 4 K-blocks x
 * (fp16 hi 16 KB, fp16 lo 16 KB) in the SWIZZLE_128B layout; ceil(n / 128) * 128 KB.                                           */
int npcd_tc_rows_to_image(const float* rows, long long n, void* image, void* stream);
int npcd_tc_image_to_rows(const void* image, long long n, float* rows, void* stream);
/* self-test: out[s,:] = x[s,:] @ W^T + b for one packed 256x256 layer; `image` = npcd_tc_rows_to_image(x) */
int npcd_tc_linear_probe(const void* image, const long long* n_rows_dev, long long capacity, const npcd_tc_layer* layer, float* out,
                         int* error_flag, int num_sms, void* stream);
/* The "f16 + e4m3 x 2" operand scheme (inference): an activation y travels as fp16(8 y) plus the two bytes e4m3((8 y - fp16(8 y)) 2^8)
 * and e4m3(y / 2); a weight w (times the layer's power-of-two `scale`) as fp16(2^13 w), e4m3(2^5 w) and e4m3((2^13 w - fp16(2^13 w)) 2^4).
 * A layer is then ONE kind::f16 product plus TWO kind::f8f6f4 products (twice the tensor rate) into the same fp32 accumulator:
 * relative error ~2^-16 per product instead of 2^-22, measured well inside the 1e-4 bar on RGB (tests/test_gpu_precision.py).
 * Same sizes and call shapes as the format-0 functions above.                                                                       */
/* development aid: CTA 0 of the inference pair kernel records clock64() at its phase boundaries into buf [64][32] int64 (NULL = off);
 * _heads: the same for the inference heads kernel (events: 2 l / 2 l + 1 = epilogue of layer l waits / has its accumulator,
 * 10 = tile done, 16 + 3 l .. 18 + 3 l = MMA issuer of layer l < 4: start / first operand there / last MMA issued) */
int npcd_debug_set_timeline(void* buf);
int npcd_debug_set_timeline_heads(void* buf);
int npcd_tc_pack_weights_f8(const float* w, int k_in, const int* perm, int k_pad /* multiple of 32 */, float scale, void* out, void* stream);
int npcd_tc_rows_to_image_f8(const float* rows, long long n, void* image, void* stream);
int npcd_tc_image_to_rows_f8(const void* image, long long n, float* rows, void* stream);
int npcd_tc_linear_probe_f8(const void* image, const long long* n_rows_dev, long long capacity, const npcd_tc_layer* layer, float* out,
                            int* error_flag, int num_sms, void* stream);

/* ---- generic fp32-accurate tensor-core GEMM (training path: forward / dgrad / wgrad of the nn.Linear layers, row B* of
 * SURVEY.md section 8; the reference runs them as fp32 cuBLAS GEMMs under autograd) ------------------------------------------
 * npcd_tc_pack_rows builds the pre-split fp16 hi/lo operand image (see npcd_tc_rows_to_image; here any K, padded to 64) of
 *   X'[r',k'] = src[r',k'] (transpose = 0, src is [rows, cols] with row stride ld) or src[k',r'] (transpose = 1),
 *   times (mask_src[same position] > 0 ? 1 : slope) if mask_src != NULL (LeakyReLU derivative), times *scale_dev (device scalar,
 *   a power of two; NULL = 1).  npcd_tc_image_bytes(rows', k') sizes it.
 * npcd_tc_gemm:  C[M,N] = act((A . B^T) * *out_scale_dev + bias),  A image [M,K], B image [N,K], N <= 256, act = LeakyReLU with
 *   slope act_slope (1 = identity); split_k > 1 reduces partial sums from `workspace` in a fixed order (deterministic).        */
int npcd_tc_image_bytes(long long rows, long long k, size_t* bytes);
int npcd_tc_pack_rows(const float* src, long long rows, int cols, long long ld, int transpose, const float* mask_src, float slope,
                      const float* scale_dev, void* image, void* stream);
int npcd_tc_gemm_workspace_bytes(int M, int split_k, size_t* bytes);
int npcd_tc_gemm(const void* a_image, const void* b_image, int M, int N, long long K, float* C, long long ldc, const float* bias,
                 const float* out_scale_dev, float act_slope, int split_k, void* workspace, size_t workspace_bytes, void* stream);

/* Weight gradient on the tensor cores, straight from ROW-major operand images (the same bytes the forward / dgrad kernels use as
 * K-major operands are consumed here as MN-major operands, so no transposed copy is built):
 *   C[m, j] (+)= *out_scale_dev * sum_{row < rows} A[row, m] * B[row, col_perm ? col_perm[j] : j],   m < a_cols, j < n_out,
 * A image [rows, a_cols], B image [rows, b_cols <= 256]; rows of A beyond `rows` (up to the next multiple of 128) must be zero
 * (npcd_tc_pack_rows zero-fills them).  rows_dev (optional, device int64) overrides `rows` (clamped to it).  row_splits CTAs per
 * 128-wide M half reduce disjoint row ranges into `workspace`; the partials are summed in a fixed order (deterministic).
 * flags: 0 (bit 0 swaps the descriptor's leading / stride offsets; descriptor probe only).
 * npcd_tc_image_colsum: out[j] (+)= *out_scale_dev * sum_rows X[row, col_perm ? col_perm[j] : j]  (bias gradients);
 *   workspace >= row_splits * 64 * ceil(cols / 64) floats.                                                                     */
int npcd_tc_wgrad_workspace_bytes(int a_cols, int row_splits, size_t* bytes);
/* Several weight-gradient problems (the layers of one backward) in ONE launch pair; each also yields its bias gradient
 * bias_out[m] (+)= *bias_scale_dev * sum_rows A[row, m] (optional) from the same pass.  workspace >= the sum of
 * npcd_tc_wgrad_workspace_bytes(a_cols, row_splits) over the problems.                                                        */
#define NPCD_WGRAD_MAX_GROUPS 8
typedef struct {
  const void* a_image;
  const void* b_image;
  int a_cols, b_cols;
  long long rows;
  const long long* rows_dev; /* optional */
  float* C;
  long long ldc;
  int n_out;
  const int* col_perm;        /* optional */
  const float* out_scale_dev; /* optional */
  float* bias_out;            /* optional [a_cols] */
  const float* bias_scale_dev; /* optional: bias_out uses this scale (the inverse of A's scale) instead of out_scale_dev */
  int accumulate;
} npcd_wgrad_problem;
int npcd_tc_wgrad_grouped(const npcd_wgrad_problem* problems, int n_problems, int row_splits, void* workspace,
                          size_t workspace_bytes, int flags, void* stream);
int npcd_tc_wgrad(const void* a_image, int a_cols, const void* b_image, int b_cols, long long rows, const long long* rows_dev,
                  float* C, long long ldc, int n_out, const int* col_perm, const float* out_scale_dev, int accumulate,
                  int row_splits, void* workspace, size_t workspace_bytes, int flags, void* stream);
int npcd_tc_image_colsum(const void* image, int cols, long long rows, const long long* rows_dev, float* out, int n_out,
                         const int* col_perm, const float* out_scale_dev, int accumulate, int row_splits, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ---- fused training path of the per-(sample, neighbour) MLP (autograd of fields/aggregators/mlp.py:69-88,119-121 and the four
 * hidden layers of `local_field`, utils/model.py:22-36; the reference gets it from torch autograd) --------------------------
 * npcd_pair_tc_train_fwd = stage 1 of npcd_field_tc_fwd (dense packing + gather + posenc + 4 layers + weighted aggregation into
 *   the operand image at the start of `workspace`; read it back with npcd_tc_image_to_rows) that additionally writes the STASH:
 *   per dense pair tile the operand images of the layer inputs X_0..X_3, the LeakyReLU sign masks of the four layer outputs, and
 *   each row's normalised weight / point index / sample index.  npcd_pair_stash_layout_for(capacity) gives the byte offsets.
 * npcd_pair_tc_bwd: d_agg [S,256] = dL/d(aggregated feature) -> d_kp_feat [n_obj*P,32] (+= with fp32 atomics, like the
 *   reference's index_add_) and the dP_l operand images (stash offsets dp[l]); w_t_packed[l] = npcd_tc_pack_weights of W_l^T
 *   ([in, out] row-major; l = 0: only the 32 feature rows, zero-padded to 256 rows), inv_scale[l] (HOST) their inverse scales;
 *   scale_dev [2] = {s, 1/s} from npcd_absmax_scale(d_agg, ..., target_exp = 6).
 *   Then, per layer l, with rows = max_tiles * 128 and rows_dev = stash + rows_dev:
 *     dW_l = npcd_tc_wgrad(A = stash + dp[l] (256 cols), B = stash + x[l] (l = 0: 112 cols in OUR column order), scale 1/s)
 *     db_l = npcd_tc_image_colsum(stash + dp[l]).                                                                             */
typedef struct {
  long long max_tiles;
  size_t x[4];     /* operand images of X_0 (2 K-blocks per tile) and X_1..X_3 (4 K-blocks per tile) */
  size_t dp[4];    /* operand images of dP_0..dP_3 (4 K-blocks per tile), written by npcd_pair_tc_bwd */
  size_t mask[4];  /* [tile][128][8] uint32 sign bits of the outputs of layers 0..3 */
  size_t wn, idx, samp; /* [tile][128] float / int32 / int32 */
  size_t rows_dev; /* int64: n_tiles * 128 */
  /* heads stage, per 128-sample tile (h_tiles = ceil(capacity / 128)) */
  long long h_tiles;
  size_t hx[6];    /* operand images of F (local_field.8 output), C1, C2, C3 (channel_net hidden), C4, H (shape_net hidden) */
  size_t hdp[6];   /* operand images of dP_c3, dP_c2, dP_c1, dP_c0, dP_s, dF, written by npcd_heads_tc_bwd */
  size_t hmask[5]; /* sign bits of H, C1, C2, C3, C4 */
  size_t g4;       /* fp32 [capacity,4]: dL/d(pre-sigmoid rgb, pre-softplus sigma), written by npcd_heads_tc_bwd */
  size_t d_agg;    /* fp32 [capacity,256]: dL/d(aggregated pair feature), written by npcd_heads_tc_bwd */
  size_t total;
} npcd_pair_stash_layout;
int npcd_pair_stash_layout_for(long long capacity, npcd_pair_stash_layout* out);
int npcd_pair_tc_train_fwd(const int* nbr_idx, const float* sample_pos, const float* kp_pos, const float* kp_feat,
                           const long long* n_samples_dev, long long capacity, const npcd_mlp_tc_weights* weights, void* workspace,
                           size_t workspace_bytes, const npcd_pair_stash_layout* layout, void* stash, size_t stash_bytes,
                           int* error_flag, int num_sms, void* stream);
/* Training forward of the WHOLE field: npcd_pair_tc_train_fwd, then the heads stage (local_field.8, shape_net, channel_net,
 * output activations -> rgbs [capacity,4]) stashing its layer inputs / sign masks too.  The aggregate operand image (input of
 * local_field.8, needed by its weight gradient) stays at the start of `workspace`: keep it until the backward is done.
 * npcd_heads_tc_bwd: d_rgbs [S,4] = dL/d(r,g,b,sigma), rgbs = the forward output -> stash d_agg (input of npcd_pair_tc_bwd), g4
 *   and the six dP images; w_t_packed / inv_scale (HOST) in order W_c3^T, W_c2^T, W_c1^T, W_c0^T, W_s0^T, W_4^T
 *   (channel_net.6,4,2,0, shape_net.0, local_field.8; entries 3 and 4 packed with ONE common scale); chan_out_w [3,256] and
 *   shape_out_w [256] are HOST pointers (they travel by value); scale_dev from npcd_absmax_scale(d_rgbs, ..., target_exp = 8).  */
int npcd_field_tc_train_fwd(const int* nbr_idx, const float* sample_pos, const float* kp_pos, const float* kp_feat,
                            const long long* n_samples_dev, long long capacity, const npcd_mlp_tc_weights* weights, void* workspace,
                            size_t workspace_bytes, const npcd_pair_stash_layout* layout, void* stash, size_t stash_bytes,
                            float* rgbs, int* error_flag, int num_sms, void* stream);
int npcd_heads_tc_bwd(const float* d_rgbs, const float* rgbs, const long long* n_samples_dev, long long capacity,
                      const npcd_pair_stash_layout* layout, void* stash, const void* const* w_t_packed, const float* inv_scale,
                      const float* chan_out_w, const float* shape_out_w, const float* scale_dev, int* error_flag, int num_sms,
                      void* stream);
/* scale_out [2] = {s, 1/s}, s a power of two with s * max|x| in [2^target_exp, 2^(target_exp+1)) (s = 1 if x == 0);
 * scratch4: 4 bytes of device scratch.                                                                                         */
int npcd_absmax_scale(const float* x, long long n, int target_exp, void* scratch4, float* scale_out, void* stream);
int npcd_pair_tc_bwd(const float* d_agg, const npcd_pair_stash_layout* layout, void* stash, const void* const* w_t_packed,
                     const float* inv_scale, const float* scale_dev, float* d_kp_feat, int* error_flag, int num_sms, void* stream);

/* ---- compositing: replaces Renderer.get_depths_from_shading_pts (renderers/renderer.py:95-110), VolumeRenderer.get_alpha
 * (renderers/volume_renderer.py:23-39), Renderer.ray_march (renderers/renderer.py:120-185).
 *   out_mask [n_sel], out_depth [n_sel] (UNCLAMPED, NaN -> +inf), out_rgb [n_sel,3]; range_scratch (8 bytes) accumulates the
 *   global slot-depth range (init_range != 0 resets it first, so several view chunks can share one range);
 *   npcd_clamp_depth then applies renderer.py:154-156 and records out_clamped [n] u8 (optional; needed by the backward).
 *   Backward: g_* may be NULL; g_rgbs [S,4] = d/d(r,g,b,sigma).                                                                */
/*   sample_t (optional): the slot depths as a dense [S] array (npcd_knn_fill), else sample_pos[.][3] is read;
 *   slot (optional, voxel-compat mode): slot index of every kept sample (npcd_voxel_slots); a sample whose successor is not in the
 *   next slot is followed by a hole and gets delta = 0 (alpha = 0), leading holes put ray_end into the clamp range.               */
int npcd_composite_fwd(const float* sample_pos, const float* sample_t, const float* rgbs, const long long* ray_offset,
                       const int* ray_ids, const float* ray_end, const unsigned char* slot, long long n_sel, int white_back,
                       float* out_mask, float* out_depth, float* out_rgb, void* range_scratch, int init_range, void* stream);
int npcd_clamp_depth(float* depth, long long n, const void* range_scratch, unsigned char* out_clamped, void* stream);
int npcd_composite_bwd(const float* sample_pos, const float* sample_t, const float* rgbs, const long long* ray_offset,
                       const unsigned char* slot, long long n_sel, int white_back, const float* g_rgb, const float* g_mask,
                       const float* g_depth, const float* out_mask, const float* out_depth, const unsigned char* clamped,
                       float* g_rgbs, void* stream);

/* ---- TV loss of the neural point cloud: the second caller of the kNN boundary (npcd/losses/neural_point_cloud_tv_loss.py:28-83).
 *   nbr_idx [n,8] = npcd_knn_points of every point against its own cloud (global indices, -1 padded);
 *   tv_out[p] = weight * sum_n ||f_n - f_p||_1 / (||x_n - x_p||_2 + 1e-5);  backward: d_feat [n,F] += d tv / d feat * g_tv[p]
 *   (fp32 atomics, like the reference's index_add_); no gradient to the positions (the loss detaches them, :39).              */
int npcd_tv_loss_fwd(const float* kp_pos, const float* kp_feat, const int* nbr_idx, long long n_points_total, int feat_dim,
                     float weight, float* tv_out, void* stream);
int npcd_tv_loss_bwd(const float* kp_pos, const float* kp_feat, const int* nbr_idx, long long n_points_total, int feat_dim,
                     float weight, const float* g_tv, float* d_feat, void* stream);

/* ---- embedding-side training step ---------------------------------------------------------------
 * npcd_embed_fwd / _bwd: VariationalEmbedding.forward + get_mean_log_var_std (npcd/models/pointnerf/embeddings/
 *   variational_embedding.py:36-70) fused: table [n_obj, P*2F] rows obj_idx [n_slots] int64 -> feats = mean + exp(0.5 lv) * eps
 *   (eps NULL: feats = mean, the eval path), mean, log_var, std, each [n_slots,P,F] (any output may be NULL).  Backward writes the
 *   COMPACT row gradient d_rows [n_slots, P*2F] (duplicated objects are summed by the optimiser kernel, like nn.Embedding's
 *   backward would) from g_feats / g_mean / g_log_var / g_std (any may be NULL).
 * npcd_kl_fwd / _bwd: NeuralPointCloudKLLoss (npcd/losses/neural_point_cloud_kl_loss.py:36-37):
 *   kld [n] = -0.5 * weight * sum_f (1 + lv - mean^2 - exp(lv)).
 * npcd_embed_adam_rows: torch.optim.Adam as the reference trainer runs it over the table (npcd/train/pointnerf_training.py:101-102,
 *   152; defaults: no weight decay, no amsgrad), applied to the rows obj_idx [n_slots] only yet EQUAL to the dense optimiser:
 *   row_step [n_obj] int32 holds the step each row was last brought up to (0 = never touched, exp_avg = exp_avg_sq = 0) and the
 *   missed zero-gradient steps are replayed before `step` (1-based) is applied with d_rows.  obj_idx == NULL and d_rows == NULL:
 *   bring rows 0..n_slots-1 up to date with `step` (before reading the table for evaluation or a checkpoint).                  */
int npcd_embed_fwd(const float* table, const long long* obj_idx, int n_slots, int n_points, int feat_dim, const float* eps,
                   float* feats, float* mean, float* log_var, float* std, void* stream);
int npcd_embed_bwd(const float* table, const long long* obj_idx, int n_slots, int n_points, int feat_dim, const float* eps,
                   const float* g_feats, const float* g_mean, const float* g_log_var, const float* g_std, float* d_rows,
                   void* stream);
int npcd_kl_fwd(const float* mean, const float* log_var, long long n_points_total, int feat_dim, float weight, float* kld,
                void* stream);
int npcd_kl_bwd(const float* mean, const float* log_var, long long n_points_total, int feat_dim, float weight, const float* g_kld,
                float* d_mean, float* d_log_var, void* stream);
int npcd_embed_adam_rows(float* table, float* exp_avg, float* exp_avg_sq, int* row_step, const long long* obj_idx, int n_slots,
                         long long row_len, const float* d_rows, int step, double lr, double beta1, double beta2, double adam_eps,
                         void* stream);

/* ---- decode post-processing of eval_diffusion (npcd/eval/diffusion_evaluation.py:169-173): channels [n_views, res*res, 3] ->
 * images [n_views, 3, res, res] (npcd/utils/util.py:199-203 unflatten_pred), optionally clipped to [0,1] and quantised to
 * round(x * 255) / 255 (round-half-to-even, as numpy).                                                                         */
int npcd_channels_to_images(const float* channels, long long n_views, int resolution, int quantize, float* images, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NPCD_B200_H */
