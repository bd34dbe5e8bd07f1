/*
 * avlmaps_b200 -- C-ABI of the B200-native AVLMaps hot paths.
 *
 * The reference (avlmaps/AVLMaps, pure Python) has no FFI; the boundary it
 * offers is the Python surface of `avlmaps.map`.  The
 * entry points below are what a ctypes stub inside those reference files
 * would bind; each one names the reference lines whose body it replaces
 * (paths relative to /root/reference).  INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain C types only; every function returns AVL_OK (0) or an AVL_ERR_*
 *     code and leaves a message for avl_last_error() (thread-local).
 *   - `flags & AVL_ON_DEVICE`: the data pointers of that call are device
 *     pointers (HBM-resident, e.g. torch tensors' data_ptr()); otherwise they
 *     are host pointers and the call performs the H2D / D2H copies itself.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *     Host-pointer calls synchronise the stream before returning; device-
 *     pointer calls only enqueue work (except where a count must be read
 *     back, which is documented per call).
 *   - inputs are borrowed for the duration of the call; handles own their
 *     device memory; outputs are caller-allocated.
 *   - threading: a handle (avl_map / avl_builder / avl_bounds) owns a per-handle
 *     workspace and must not be used from two threads at once; different handles
 *     may be used concurrently.  The reference is single-threaded and synchronous.
 *   - there is NO CPU implementation behind any entry point: without a
 *     CUDA device every compute call fails with AVL_ERR_CUDA.
 */
#ifndef AVLMAPS_B200_H_
#define AVLMAPS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVL_OK 0
#define AVL_ERR_CUDA 1        /* CUDA runtime/driver error (message has the cudaError string) */
#define AVL_ERR_ARG 2         /* invalid argument */
#define AVL_ERR_UNSUPPORTED 3 /* shape outside what the kernels support */
#define AVL_ERR_STATE 4       /* call order violated (e.g. export before finalize) */

#define AVL_ON_DEVICE 1 /* data pointers of this call are device pointers */
#define AVL_DEPTH_U16_MM 2 /* avl_frame.depth points at (h, w) uint16 millimetres; the kernel evaluates
                              float64(d) / 1000.0 like `load_depth_img(...) / 1000.0`
                              (vlmap_builder_multi_floor.py:103,128) */

#define AVL_ASYNC 16 /* avl_sim_topk with HOST pointers: enqueue the H2D copy of the queries, the kernels and the D2H copy of
                         the result on `stream` and return WITHOUT synchronising.  The host buffers must be pinned and must
                         not be touched until the stream has passed this call (cudaStreamSynchronize / an event): the way
                         a serving loop keeps several query batches in flight. */

#define AVL_PIPELINED 32 /* avl_sim_topk (stats == NULL): the tail of the call -- exact re-score of the survivors, the
                            device-side fallback and, with host pointers, the copy of the result -- is enqueued on the
                            map's own tail stream, where it runs next to the FOLLOWING call's screen kernel (a serving
                            loop hides ~50 us of a ~0.96 ms step that way).  The results of a call are ordered on `stream`
                            only after the NEXT pipelined call on the map, or after avl_map_flush(map, stream); query,
                            scale and output buffers of a call must stay untouched until then, so consecutive calls take
                            distinct buffers.  Work that should follow a call's results directly (the peer exchange of a
                            slab-sharded map) can be enqueued on avl_map_tail_stream instead. */

#define AVL_MAP_F16 4 /* avl_map_create: keep the tensor-core copy of the map (and of the queries) in fp16 instead of
                         bf16.  Same tcgen05 kind::f16 rate and bytes; the rounding residual, hence the rigorous error
                         band that decides what is re-scored exactly, is 8x smaller.  Results are identical either
                         way.  Falls back to bf16 by itself when a value exceeds the fp16 range. */

#define AVL_FEAT_F16 8 /* avl_builder_add_frame(s): avl_frame.feat points at fp16 values, in either layout.  LSeg emits
                          `logit_scale * normalize(x).half()` (avlmaps/lseg/modules/models/lseg_net.py:318-321), so the
                          float32 array get_lseg_feat returns holds fp16-exact values; handing the halves over directly
                          halves the bytes of the hand-off (415 -> 208 MB per 390x520x512 frame) and changes no result.
                          AVL_FEAT_HWC | AVL_FEAT_F16 is the device-side hand-off of an encoder that stays on the GPU:
                          pixel-major fp16 rows, read as they are by the scatter kernel (no transposition, half the
                          feature bytes).  bf16 is not offered: it would round LSeg's fp16 values. */

#define AVL_MAX_QUERIES 256 /* per call; larger batches are chunked by the host layer */
#define AVL_MAX_TOPK 128

/* combine rules of avl_fuse_topk (habitat_lang_robot.py:223, avlmap.py:92-97,124-131; paper: product) */
#define AVL_FUSE_PRODUCT 0
#define AVL_FUSE_MAX 1
#define AVL_FUSE_SUM 2

/* feature layouts accepted by avl_builder_add_frame */
#define AVL_FEAT_CHW 0 /* (1, D, FH, FW) fp32, what get_lseg_feat returns (lseg_utils.py:101-102) */
#define AVL_FEAT_HWC 1 /* (FH, FW, D) fp32 pixel-major, the B200-friendly hand-off */

typedef struct avl_map avl_map;         /* device-resident voxel feature map (grid_feat) */
typedef struct avl_builder avl_builder; /* device-resident map under construction */
typedef struct avl_bounds avl_bounds;   /* running min / max of back-projected points (multi-floor pass 1) */

/* Filled by the index calls when a non-NULL pointer is passed. */
typedef struct avl_index_stats {
  int64_t n_rows;
  int32_t dim;
  int32_t n_queries;
  int32_t cta_group;          /* tcgen05 variant that ran: 1 / 2 = cta_group of sim_screen.cu, 0 = no tensor-core kernel */
  int32_t n_launches;         /* kernels launched by the call */
  int64_t n_flagged;          /* argmax: rows whose bf16 margin was inside the error band (re-ranked exactly) */
  int64_t n_candidates;       /* top-k: (row, query) pairs that passed the screen threshold */
  int32_t n_fallback_queries; /* top-k: queries whose candidate buckets overflowed -> re-scored exactly (on the device) */
  int32_t sample_rows;        /* top-k: rows used for the threshold estimate */
  float ms_screen;            /* CUDA-event time of the main tcgen05 launch (avl_set_profiling(1)), else 0 */
  float ms_total;             /* CUDA-event time of the whole call's device work, else 0 */
} avl_index_stats;

/* ---- library -------------------------------------------------------------------------- */
int avl_version(void);              /* 100*major + minor; no CUDA call */
const char* avl_last_error(void);   /* thread-local message of the last failing call */
int avl_device_count(int* count);   /* number of CUDA devices (0 on a CPU box) */
int avl_set_device(int device);     /* cudaSetDevice for the calling thread */
int avl_set_profiling(int enabled); /* record CUDA events inside index calls (read through `stats` or
                                       avl_map_screen_times) */

/* ---- landmark-index path ------------------------------------------------------------- */

/* Upload (or adopt a device copy of) grid_feat (n, dim) fp32 C-contiguous, build the bf16
 * tensor-core copy and the per-row norms / rounding residuals.
 * Replaces the host-RAM residency of VLMap.grid_feat after load_3d_map
 * (avlmaps/map/vlmap.py:50-65, avlmaps/utils/mapping_utils.py:508-541). */
int avl_map_create(const float* grid_feat, int64_t n, int32_t dim, int flags, void* stream, avl_map** out);
int avl_map_destroy(avl_map* map);
int avl_map_shape(const avl_map* map, int64_t* n, int32_t* dim);
int avl_map_operand_f16(const avl_map* map); /* 1 if the tensor-core copy is fp16, 0 if bf16 */
int64_t avl_map_device_bytes(const avl_map* map);

/* scores[i, q] = scale[q] * inv_norm[i] * <grid_feat[i], queries[q]>   (fp32, (n, nq) C-contiguous)
 * inv_norm = 1 unless normalize_map; scale = 1 if NULL.  Accumulated in fp64, rounded once.
 * Replaces `scores_list = map_feats @ text_feats.T` (avlmaps/utils/clip_utils.py:227-229,236-240;
 * duplicate avlmaps/utils/index_utils.py:93-106) and the scaled audio form
 * (avlmaps/map/sound_map.py:108-109). */
int avl_sim_dense(avl_map* map, const float* queries, int32_t nq, const float* scale, int normalize_map,
                  float* out_scores, int flags, void* stream);

/* out_argmax[i] = argmax_q scores[i, q], ties -> lowest q (numpy argmax).  int32 (n,).
 * Replaces get_lseg_score + `np.argmax(scores_mat, axis=1)` (avlmaps/map/vlmap.py:113-124,
 * avlmaps/utils/index_utils.py:153-161, avlmaps/robot/habitat_lang_robot.py:243).
 * tcgen05 bf16 screen; rows whose top-2 margin is inside the rigorous bf16 error band are
 * re-scored exactly, so the result equals the argmax of avl_sim_dense's scores.
 * Device-pointer calls read one counter back (stream sync). */
int avl_sim_argmax(avl_map* map, const float* queries, int32_t nq, const float* scale, int normalize_map,
                   int32_t* out_argmax, int flags, void* stream, avl_index_stats* stats);

/* Per query the k best rows: out_idx (nq, k) int64, out_score (nq, k) fp32, sorted by
 * (score desc, row asc); slots past n are idx -1 / score -inf.  Scores are the exact ones.
 * Generalises `grid_pos[np.argmax(heat)]` (avlmaps/robot/habitat_lang_robot.py:427-430) and the
 * argsort retrieval precedent (avlmaps/utils/clip_utils.py:86-93) to k > 1.
 * With AVL_ON_DEVICE and stats == NULL the call is ASYNCHRONOUS: it enqueues its kernels on `stream` and returns;
 * the results are ordered on the stream like any kernel's output (the exact fallback for queries whose candidate
 * buckets overflow is decided on the device).  Host pointers, or stats != NULL, synchronise the stream before
 * returning.  Calls on one map must be issued from one thread and one stream at a time (they share its workspace). */
int avl_sim_topk(avl_map* map, const float* queries, int32_t nq, const float* scale, int normalize_map,
                 int32_t k, int64_t* out_idx, float* out_score, int flags, void* stream,
                 avl_index_stats* stats);

/* Pipelined calls (AVL_PIPELINED): make `stream` wait for the tails of all earlier calls on the map -- their results
 * are valid for work enqueued on `stream` afterwards.  avl_map_tail_stream: the cudaStream_t those tails run on. */
int avl_map_flush(avl_map* map, void* stream);
int avl_map_tail_stream(avl_map* map, void** out_stream);

/* With avl_set_profiling(1) every avl_sim_topk call records a CUDA-event pair around its main screen launch into a
 * ring of 256; this reads the times (ms, oldest first) of the calls since the last read into out_ms[0..cap) and
 * clears the ring.  Synchronises the last recorded event.  *n_out = number written. */
int avl_map_screen_times(avl_map* map, float* out_ms, int32_t cap, int32_t* n_out);

/* Diagnostic: the raw bf16 tensor-core scores (n, nq) fp32 of the screen kernel, no correction.
 * cta_group = 1 or 2 selects the tcgen05 variant (0 = engine's choice). */
int avl_sim_screen_dense(avl_map* map, const float* queries, int32_t nq, int32_t cta_group,
                         float* out_scores, int flags, void* stream);

/* Exact top-k of a vector: (value desc, index asc).  Serves get_max_pos_3d on any fused heat. */
int avl_topk_f32(const float* values, int64_t n, int32_t k, int64_t* out_idx, float* out_val, int flags,
                 void* stream);

/* Multi-GPU: merge per-slab top-k results (n_shards, nq, k) carrying GLOBAL row ids (-1 = empty) into the
 * global (nq, k), same (score desc, row asc) order.  Device pointers only: the input is the buffer the
 * one NCCL all-gather of the path filled.  n_shards * k <= 1024. */
int avl_merge_topk(const int64_t* idx, const float* val, int32_t n_shards, int32_t nq, int32_t k,
                   int64_t* out_idx, float* out_val, int flags, void* stream);

/* Cross-modal goal selection (BASELINE config 3).  For pair j < n_pairs:
 *   h_a = minmax_i(score_a[:, j]), h_b = minmax_i(score_b[:, j])   (sound_map.py:151-152,
 *   habitat_lang_robot.py:213-214), heat = combine(h_a, h_b), result = top-k(heat).
 * map_a / map_b must have the same number of rows. out_idx/out_heat are (n_pairs, k). */
int avl_fuse_topk(avl_map* map_a, const float* queries_a, const float* scale_a, int normalize_a,
                  avl_map* map_b, const float* queries_b, const float* scale_b, int normalize_b,
                  int32_t n_pairs, int32_t combine, int32_t k, int64_t* out_idx, float* out_heat,
                  int flags, void* stream);

/* heat[i] = 1 on target voxels (mask[i] != 0), else clip(1 - min_t ||pos_i - pos_t|| / cell_size * decay_rate, 0, 1).
 * grid_pos (n, 3) int32, mask (n,) uint8 (numpy bool), out_heat (n,) fp32.  Bit-exact restatement of
 * get_heatmap_from_mask_3d (avlmaps/utils/visualize_utils.py:29-49; twin habitat_lang_robot.py:242-265),
 * the O(N_other * N_target) Python loop behind AVLMap.index_object (avlmaps/map/avlmap.py:67-76).
 * An empty mask is an argument error (the reference's np.argmin raises). */
int avl_heat_from_mask_3d(const int32_t* grid_pos, const uint8_t* mask, int64_t n, double cell_size,
                          double decay_rate, float* out_heat, int flags, void* stream);

/* 2-D heat from point sources on the (rows, cols) top-down grid: replaces the per-frame / per-segment
 * full-grid distance_transform_edt loops of AVLMap.index_area_2d (avlmaps/map/avlmap.py:78-98; mode 0:
 * max-combine of clip(s - decay*dist, 0, 1), float64 out) and AVLMap.index_sound_2d (avlmap.py:111-133;
 * mode 1: float32 running sum of max(con - con*dist*decay, 0) in segment order, float32 out), BEFORE the
 * final min-max.  cells (n_src, 2) int32 (row, col) sorted by group; group_start (n_groups + 1,) int32
 * offsets (an empty group is a frame that fell outside the grid); conf (n_groups,) fp32.  Host pointers. */
int avl_heat2d_sources(const int32_t* cells, const int32_t* group_start, const float* conf, int32_t n_groups,
                       int32_t rows, int32_t cols, double decay_rate, int32_t mode, void* out_heat, int flags,
                       void* stream);

/* The tail of index_area(_2d) / index_sound(_2d): min-max normalise the 2-D map IN PLACE in its own dtype
 * (`(dist_map - min) / (max - min)`, avlmaps/map/avlmap.py:97 float64, :131 float32) when `normalize`, and lift it
 * to the voxels, heat3d[i] = float32(heat2d[grid_pos[i, 0], grid_pos[i, 1]]) (avlmap.py:100-109, 135-144: the Python
 * loop over np.where(occupied_ids != -1)).  n == 0 skips the lift.  Host pointers. */
int avl_heat2d_normalize_lift(void* heat2d, int32_t is_f64, int32_t rows, int32_t cols, int32_t normalize,
                              const int32_t* grid_pos, int64_t n, float* out_heat3d, int flags, void* stream);

/* Image modality after localisation: out[i] = clip(con - decay_rate * ||grid_pos[i, :2] - (row, col)||, 0, 1),
 * float64 (n,).  Replaces the numpy lines of AVLMap.index_image (avlmaps/map/avlmap.py:156-162); the
 * localisation itself (HLoc, visual_map.py) is out of scope. */
int avl_heat_planar(const int32_t* grid_pos, int64_t n, double row, double col, double con, double decay_rate,
                    double* out_heat, int flags, void* stream);

/* ---- map-build path ------------------------------------------------------------------ */

typedef struct avl_grid_spec {
  int32_t gs;       /* grid_size: rows = cols = gs            (vlmap_builder.py:62)  */
  int32_t vh;       /* int(camera_height / cs)                (vlmap_builder.py:201) */
  double cs;        /* cell_size in metres                    (vlmap_builder.py:61)  */
  int32_t dim;      /* feature dim D                          (vlmap_builder.py:202) */
  int64_t capacity; /* initial voxel rows (reference: gs*gs, doubled on demand :151-152,286-311) */
} avl_grid_spec;

typedef struct avl_frame {
  const float* depth; /* (h, w) fp32 metres (mapping_utils.py:231); uint16 mm with AVL_DEPTH_U16_MM */
  int32_t h, w;
  const float* feat; /* per-pixel features, layout below (vlmap_builder.py:123-126)             */
  int32_t fh, fw;
  int32_t feat_layout;       /* AVL_FEAT_CHW | AVL_FEAT_HWC                                     */
  const uint8_t* rgb;        /* (h, w, 3) u8 or NULL (vlmap_builder.py:118-119)                 */
  const int32_t* sample_idx; /* pixel ids v*w+u in the order `shuffle_mask[::rate]` yields
                                (vlmap_builder.py:275-277); NULL = every pixel in raster order  */
  int32_t n_samples;
  double kinv[9];  /* np.linalg.inv(cam_calib_mat), row-major (mapping_utils.py:237)            */
  double k[9];     /* cam_calib_mat (vlmap_builder.py:98,141)                                   */
  double kfeat[9]; /* get_sim_cam_mat(fh, fw) (mapping_utils.py:591-596)                        */
  double tf[16];   /* pc_transform = tf @ base_transform @ base2cam_tf (vlmap_builder.py:133)   */
  double min_depth, max_depth; /* 0.1, 6 at the reference call site (vlmap_builder.py:129)      */
} avl_frame;

int avl_builder_create(const avl_grid_spec* spec, avl_builder** out);
int avl_builder_destroy(avl_builder* b);

/* Back-project one frame and fuse it (vlmap_builder.py:129-178): fp64 geometry without FMA
 * contraction, first-touch voxel ids in (frame, sample) order, alpha-weighted accumulation.
 * Frames must be added in the reference's frame order. */
int avl_builder_add_frame(avl_builder* b, const avl_frame* frame, int flags, void* stream);

/* Several consecutive frames in one call.  With device pointers and pixel-major features the samples of up to 16
 * frames are concatenated into ONE geometry / id-scan / scatter launch triple (the first-touch key already orders
 * them by frame, then sample), which amortises the launch and tail cost of the small kernels; the result is
 * identical to n_frames avl_builder_add_frame calls.  Host pointers or AVL_FEAT_CHW fall back to that loop. */
int avl_builder_add_frames(avl_builder* b, const avl_frame* frames, int32_t n_frames, int flags, void* stream);

/* Slab-sharded build: count n_frames frames that were NOT handed to this builder because the caller has shown that
 * none of their points can fall into its row slab (avlmaps_b200.sharded.frame_row_range: the rows the camera frustum
 * can reach).  The first-touch order is (frame, sample) over ALL frames of the build, so the skipped frames must
 * keep their frame numbers. */
int avl_builder_skip_frames(avl_builder* b, int32_t n_frames);

/* Allocate the per-sample scratch of a launch triple for calls of up to samples_per_call samples (the sum over the
 * frames of one avl_builder_add_frames call) ahead of time.  Without it the scratch grows -- cudaFree + cudaMalloc,
 * which synchronise the device -- inside the first call that needs more, i.e. somewhere in the frame loop of a slab
 * build whose first frames are skipped.  No reference counterpart (the reference allocates per frame in numpy). */
int avl_builder_reserve(avl_builder* b, int64_t samples_per_call, void* stream);

/* voxels created so far (max_id, vlmap_builder.py:164-170); synchronises the stream. */
int avl_builder_num_voxels(avl_builder* b, int64_t* n, void* stream);
/* points that passed the depth / grid / feature-bounds tests so far (P_acc of SURVEY 8d). */
int avl_builder_num_accepted(avl_builder* b, int64_t* n, void* stream);
/* bytes the builder has uploaded from HOST pointers so far (depth, sample list, rgb and features; with host-resident
 * channel-major features only the pixel rows the accepted points read are uploaded) */
int avl_builder_h2d_bytes(avl_builder* b, int64_t* n);

/* Export arrays[:max_id] + occupied_ids like _save_3d_map (vlmap_builder.py:313-327):
 * grid_feat (V, D) f32, grid_pos (V, 3) i32, weight (V,) f32, occupied_ids (gs, gs, vh) i32,
 * grid_rgb (V, 3) u8 (may be NULL).  Any pointer may be NULL to skip that array.  Can be
 * called repeatedly (the reference saves every 100 frames, :181-183). */
int avl_builder_export(avl_builder* b, float* grid_feat, int32_t* grid_pos, float* weight,
                       int32_t* occupied_ids, uint8_t* grid_rgb, int flags, void* stream);

/* Resume: adopt a saved map as the builder's state, the way _init_map reloads vlmaps.h5df before the
 * frame loop (vlmap_builder.py:212-222; max_id = grid_feat.shape[0]) -- every frame added afterwards is
 * fused ON TOP of it, exactly like the reference (whose loop never consults mapped_iter_set, :102-180):
 * a reloaded voxel continues its running mean from (grid_feat, weight), new cells get ids from n_voxels on.
 * occupied_ids is rebuilt from grid_pos.  Only valid on a fresh builder (no frame added yet).
 * grid_rgb may be NULL. */
int avl_builder_import(avl_builder* b, const float* grid_feat, const int32_t* grid_pos, const float* weight,
                       const uint8_t* grid_rgb, int64_t n_voxels, int flags, void* stream);

/* Hand the finished map to the index path without leaving HBM. */
int avl_builder_to_map(avl_builder* b, void* stream, avl_map** out);

/* ---- multi-floor (global-frame) build: VLMapBuilderMultiFloor.create_global_map ---------
 * (avlmaps/map/vlmap_builder_multi_floor.py:60-199).  Same fusion rule and first-touch ids as
 * the mobile-base build; what differs is the cell function
 *     row, height, col = np.round((p_global - pcd_min) / cs).astype(int)           (:146)
 * (round-half-even, x -> row, y -> height, z -> col), the acceptance test (only `row >= n_row or
 * col >= n_col` is rejected, :151-153; NEGATIVE indices wrap like numpy's, and grid_pos keeps the
 * unwrapped values, :176), and the depth source (uint16 mm PNG / 1000.0, max_depth 100).  Where the
 * reference would die with an IndexError (height >= n_height, any index < -size) the point is
 * rejected and counted (avl_builder_num_rejected_oob). */
typedef struct avl_global_grid_spec {
  int32_t n_row, n_col, n_height; /* grid_size[[0, 2, 1]], grid_size = ceil((pcd_max - pcd_min) / cs + 1) (:222-224) */
  double cs;
  double pcd_min[3];              /* (x, y, z) minimum of the pass-1 cloud (:117)                   */
  int32_t dim;
  int64_t capacity;               /* initial voxel rows; 0 = n_row * n_col (:225)                   */
} avl_global_grid_spec;

int avl_builder_create_global(const avl_global_grid_spec* spec, avl_builder** out);
int avl_builder_num_rejected_oob(avl_builder* b, int64_t* n, void* stream);

/* Pass 1 of create_global_map (:97-118): running component-wise min / max of
 * transform_pc(depth2pc(depth)[:, sample_idx][:, mask], tf) over the frames added.  Only depth, h, w,
 * sample_idx, n_samples, kinv, tf, min_depth, max_depth of the frame are read.  min / max are exact
 * (order-free), so pcd_min / pcd_max equal the reference's bit for bit.
 * avl_bounds_get synchronises; n_points == 0 leaves +inf / -inf (the reference's np.min raises). */
int avl_bounds_create(avl_bounds** out);
int avl_bounds_destroy(avl_bounds* b);
int avl_bounds_add_frame(avl_bounds* b, const avl_frame* frame, int flags, void* stream);
int avl_bounds_get(avl_bounds* b, double pcd_min[3], double pcd_max[3], int64_t* n_points, void* stream);

/* ---- slab-sharded build (one process per GPU) ----------------------
 * A builder restricted to the grid rows [row_lo, row_hi): every rank sees every frame, runs the
 * (cheap) geometry for every sample and keeps the points of its own slab; there is no collective in
 * the frame loop.  Local voxel ids are first-touch order WITHIN the slab; the global first-touch id
 * of a voxel is the rank of its first-touch key (frame_seq << 32 | sample position) among the keys
 * of all slabs: gather avl_builder_export_keys of every rank (one all-gather at finalize) and call
 * avl_rank_keys.  Must be set before the first frame. */
int avl_builder_set_slab(avl_builder* b, int32_t row_lo, int32_t row_hi);
/* keys (V,) uint64 of the builder's voxels, ascending (voxel ids are assigned in key order). */
int avl_builder_export_keys(avl_builder* b, uint64_t* keys, int flags, void* stream);
/* keys_all: concatenation of n_shards ascending key lists, shard s at [offsets[s], offsets[s+1]).
 * out_global_ids[j] (offsets[shard+1]-offsets[shard] entries, int64) = number of keys of ALL shards
 * smaller than the j-th key of `shard` = its voxel id in a single-GPU build.  offsets: host pointer. */
int avl_rank_keys(const uint64_t* keys_all, const int64_t* offsets, int32_t n_shards, int32_t shard,
                  int64_t* out_global_ids, int flags, void* stream);

/* ---- slab-sharded index: fused exchange + merge over NVLink peer memory ----
 * One object per rank (one process per GPU).  Every rank creates it, publishes the CUDA-IPC handle of its receive
 * buffer (avl_p2p_local_handle, avl_p2p_handle_bytes() bytes), gathers the handles of all ranks through any host
 * channel (torch.distributed.all_gather_object in avlmaps_b200/sharded.py) and maps them (avl_p2p_connect).
 * avl_p2p_exchange_merge then takes this slab's (nq, k) result (device pointers, as avl_sim_topk left it in HBM;
 * ids are made global inside the kernel: + row_offset for a contiguous slab, or a lookup in global_ids (device
 * pointer, one int64 per slab row) for the slab of a sharded build), stores it into every peer's buffer with plain
 * peer stores, and merges what the peers delivered -- ONE kernel per query batch, no host round trip, no NCCL call.
 * Every rank must call it for every batch with the same nq and k, in the same order (the epoch counter is per object);
 * calls on one object go to one stream at a time.  out_idx / out_val (nq, k): the global top-k, (score desc, row asc),
 * on every rank.  A peer that never delivers trips a ~10 s in-kernel watchdog instead of hanging the GPU: the queries
 * it was missing for come back EMPTY (idx -1, score -inf), never merged from stale bytes, and every later call on the
 * object fails with AVL_ERR_STATE (avl_p2p_status: the source rank that never arrived, or -1).
 * Default exchange of avlmaps_b200.sharded.ShardedMap (AVL_P2P_EXCHANGE=0 selects the NCCL all-gather +
 * avl_merge_topk form); both are compared bit for bit on 2 / 4 / 8 GPUs by tools/sharded_index_check.py. */
typedef struct avl_p2p avl_p2p;
int avl_p2p_create(int32_t rank, int32_t world, int32_t nq_max, int32_t k_max, avl_p2p** out);
int avl_p2p_handle_bytes(void);
int avl_p2p_local_handle(avl_p2p* p, uint8_t* handle);
int avl_p2p_connect(avl_p2p* p, const uint8_t* handles /* world * avl_p2p_handle_bytes() */);
int avl_p2p_exchange_merge(avl_p2p* p, const int64_t* idx, const float* val, int32_t nq, int32_t k,
                           int64_t row_offset /* added to every id >= 0: slab-local -> global rows */,
                           const int64_t* global_ids /* or NULL; replaces row_offset: id -> global_ids[id] */,
                           int64_t* out_idx, float* out_val, int flags, void* stream);
int avl_p2p_status(avl_p2p* p, int32_t* timed_out_source, void* stream);
int avl_p2p_destroy(avl_p2p* p);

#ifdef __cplusplus
}
#endif
#endif /* AVLMAPS_B200_H_ */
