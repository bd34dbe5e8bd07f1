/*
 * ORACLE -- test infrastructure only.  Never imported, linked or executed by the product path
 * (avlmaps_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * legs may use it, and only as the checker / the timed CPU baseline.
 *
 * Plain-C restatement of the reference's sequential map-build loop,
 *   /root/reference/avlmaps/map/vlmap_builder.py:129-178   (per-frame body of create_mobile_base_map)
 * with the helpers it calls:
 *   depth2pc              avlmaps/utils/mapping_utils.py:226-251
 *   transform_pc          avlmaps/utils/mapping_utils.py:305-315
 *   base_pos2grid_id_3d   avlmaps/utils/mapping_utils.py:345-349
 *   project_point         avlmaps/utils/mapping_utils.py:599-605
 * Arithmetic follows numpy >= 2 (NEP 50) as executed in this container: geometry in float64; the
 * matrix products accumulate left to right with explicit fma() like OpenBLAS (see dot3 below), every
 * element-wise operation rounds separately (compile with -ffp-contract=off); the fusion update computes
 * fl32(g*w) in float32, everything else of the update in float64, and rounds once on the store
 * into the float32 arrays.  Pinned against the reference itself by tests/golden (see
 * tests/golden/gen_golden.py, which runs the real VLMapBuilder through oracle/ref_shim.py).
 *
 * The multi-floor builder (avlmaps/map/vlmap_builder_multi_floor.py:60-199) shares the loop; its cell
 * function (np.round of (p - pcd_min)/cs as row, height, col, :146), its acceptance test (:151-153,
 * numpy negative-index wrap-around) and its depth source (uint16 mm / 1000.0, :103,128) are the
 * `mode == 1` branches below; oracle_frame_bounds restates the first pass (:97-118).
 *
 * Not restated (documented in DESIGN.md): the dtype drift after _reserve_map_space
 * (vlmap_builder.py:286-311) and the dead height_map / cv_map outputs (:145-147).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int32_t gs, vh, dim;   /* mode 0: occupied_ids is (gs, gs, vh) */
  int32_t n0, n1, n2;    /* occupied_ids dims (rows, cols, heights) in either mode */
  int32_t mode;          /* 0 mobile-base grid, 1 global-frame (multi-floor) grid */
  double origin[3];      /* pcd_min */
  int64_t n_oob;         /* mode 1: points on which the reference would raise IndexError */
  double cs;
  int64_t capacity;
  int64_t max_id;
  int64_t n_accepted;
  float* grid_feat;      /* (capacity, dim) */
  int32_t* grid_pos;     /* (capacity, 3)   */
  float* weight;         /* (capacity,)     */
  int32_t* occupied_ids; /* (gs, gs, vh), -1 = empty */
  uint8_t* grid_rgb;     /* (capacity, 3)   */
} oracle_builder;

oracle_builder* oracle_builder_create(int32_t gs, int32_t vh, double cs, int32_t dim, int64_t capacity) {
  oracle_builder* b = (oracle_builder*)calloc(1, sizeof(oracle_builder));
  b->gs = gs; b->vh = vh; b->cs = cs; b->dim = dim; b->capacity = capacity;
  b->n0 = gs; b->n1 = gs; b->n2 = vh; b->mode = 0;
  b->grid_feat = (float*)calloc((size_t)capacity * dim, sizeof(float));      /* vlmap_builder.py:202 */
  b->grid_pos = (int32_t*)calloc((size_t)capacity * 3, sizeof(int32_t));     /* :203 */
  b->weight = (float*)calloc((size_t)capacity, sizeof(float));               /* :205 */
  b->grid_rgb = (uint8_t*)calloc((size_t)capacity * 3, 1);                   /* :206 */
  size_t cells = (size_t)gs * gs * vh;
  b->occupied_ids = (int32_t*)malloc(cells * sizeof(int32_t));               /* :204 */
  for (size_t i = 0; i < cells; ++i) b->occupied_ids[i] = -1;
  return b;
}

/* VLMapBuilderMultiFloor._init_map (vlmap_builder_multi_floor.py:217-241): occupied_ids is
 * grid_size[[0, 2, 1]] = (n_row, n_col, n_height), row capacity grid_size[0] * grid_size[2] */
oracle_builder* oracle_builder_create_global(int32_t n_row, int32_t n_col, int32_t n_height, double cs,
                                             const double* pcd_min, int32_t dim, int64_t capacity) {
  oracle_builder* b = (oracle_builder*)calloc(1, sizeof(oracle_builder));
  b->gs = n_row; b->vh = n_height; b->cs = cs; b->dim = dim; b->capacity = capacity;
  b->n0 = n_row; b->n1 = n_col; b->n2 = n_height; b->mode = 1;
  memcpy(b->origin, pcd_min, 3 * sizeof(double));
  b->grid_feat = (float*)calloc((size_t)capacity * dim, sizeof(float));
  b->grid_pos = (int32_t*)calloc((size_t)capacity * 3, sizeof(int32_t));
  b->weight = (float*)calloc((size_t)capacity, sizeof(float));
  b->grid_rgb = (uint8_t*)calloc((size_t)capacity * 3, 1);
  size_t cells = (size_t)n_row * n_col * n_height;
  b->occupied_ids = (int32_t*)malloc(cells * sizeof(int32_t));
  for (size_t i = 0; i < cells; ++i) b->occupied_ids[i] = -1;
  return b;
}
int64_t oracle_builder_num_oob(const oracle_builder* b) { return b->n_oob; }

/* _init_map's reload of a saved map (vlmap_builder.py:212-222): the arrays ARE the loaded ones and
 * max_id = grid_feat.shape[0]; the frame loop then fuses on top.  (The reference's arrays are exactly V rows
 * long at this point and double on the next insert; the oracle just needs capacity >= V + new voxels.) */
int oracle_builder_import(oracle_builder* b, const float* grid_feat, const int32_t* grid_pos, const float* weight,
                          const uint8_t* grid_rgb, const int32_t* occupied_ids, int64_t v) {
  if (v > b->capacity || b->max_id != 0) return -1;
  memcpy(b->grid_feat, grid_feat, (size_t)v * b->dim * sizeof(float));
  memcpy(b->grid_pos, grid_pos, (size_t)v * 3 * sizeof(int32_t));
  memcpy(b->weight, weight, (size_t)v * sizeof(float));
  if (grid_rgb) memcpy(b->grid_rgb, grid_rgb, (size_t)v * 3);
  memcpy(b->occupied_ids, occupied_ids, (size_t)b->n0 * b->n1 * b->n2 * sizeof(int32_t));
  b->max_id = v;
  return 0;
}

void oracle_builder_destroy(oracle_builder* b) {
  if (!b) return;
  free(b->grid_feat); free(b->grid_pos); free(b->weight); free(b->occupied_ids); free(b->grid_rgb); free(b);
}

int64_t oracle_builder_num_voxels(const oracle_builder* b) { return b->max_id; }
int64_t oracle_builder_num_accepted(const oracle_builder* b) { return b->n_accepted; }
float* oracle_builder_grid_feat(oracle_builder* b) { return b->grid_feat; }
int32_t* oracle_builder_grid_pos(oracle_builder* b) { return b->grid_pos; }
float* oracle_builder_weight(oracle_builder* b) { return b->weight; }
int32_t* oracle_builder_occupied_ids(oracle_builder* b) { return b->occupied_ids; }
uint8_t* oracle_builder_grid_rgb(oracle_builder* b) { return b->grid_rgb; }

/* The three matrix products of the path -- cam_mat_inv @ p_2d (mapping_utils.py:245), pose @ pc_homo
 * (:314) and cam_mat @ p (:600) -- go through numpy matmul -> OpenBLAS dgemm/dgemv, whose x86-64
 * kernels accumulate over k in ascending order with fused multiply-adds:
 *     acc = m0*x ; acc = fma(m1, y, acc) ; acc = fma(m2, z, acc) [; acc = fma(m3, 1, acc)]
 * Probed in the build container on 3x3 @ 3xN and 4x4 @ 4xN for N in 1..3072: this form reproduces
 * numpy on every element, the unfused form on ~65 % (tools/probe_matmul_fma.py).  With float32
 * depths and the dataset's calibration matrices most products are exact and the two forms agree
 * (every mobile-base golden passes with either); with the multi-floor builder's uint16/1000.0
 * depths they do not, and the pixel / cell truncations sit exactly on integer boundaries. */
static inline double dot3(const double* m, double x, double y, double z) {
  return fma(m[2], z, fma(m[1], y, m[0] * x));
}
static inline double dot4h(const double* m, double x, double y, double z) { /* row . [x, y, z, 1] */
  return fma(m[3], 1.0, fma(m[2], z, fma(m[1], y, m[0] * x)));
}

/* python int(): truncation toward zero; values far outside int range are clamped (they are
 * rejected by the range checks that follow anyway) */
static inline int64_t trunc_i64(double v) {
  if (!(v > -9.0e15)) return INT64_MIN / 4;
  if (!(v < 9.0e15)) return INT64_MAX / 4;
  return (int64_t)v;
}
/* np.round(x).astype(int): round half to even (rint in the default rounding mode) */
static inline int64_t round_i64(double v) {
  if (!(v > -9.0e15)) return INT64_MIN / 4;
  if (!(v < 9.0e15)) return INT64_MAX / 4;
  return (int64_t)rint(v);
}
/* depth in metres as float64: float32 .npy widened (mode 0) or uint16 mm / 1000.0 (multi-floor :103) */
static inline double depth_at(const void* depth, int depth_u16, int32_t pix) {
  return depth_u16 ? (double)((const uint16_t*)depth)[pix] / 1000.0 : (double)((const float*)depth)[pix];
}

/* One frame.  feat is (1, D, FH, FW) float32 as get_lseg_feat returns it (lseg_utils.py:101-102).
 * sample_idx: pixel ids in the order shuffle_mask[::rate] yields (vlmap_builder.py:275-277);
 * NULL = all pixels in raster order.  Returns -1 if the capacity would be exceeded. */
static int add_frame_impl(oracle_builder* b, const void* depth, int depth_u16, int32_t h, int32_t w, const float* feat,
                          int32_t fh, int32_t fw, const uint8_t* rgb, const int32_t* sample_idx,
                          int32_t n_samples, const double* kinv, const double* k, const double* kfeat,
                          const double* tf, double min_depth, double max_depth) {
  const int32_t gs = b->gs, vh = b->vh, dim = b->dim;
  const double cs = b->cs;
  const size_t plane = (size_t)fh * fw;
  const double half = gs / 2.0;
  for (int32_t j = 0; j < n_samples; ++j) {
    const int32_t pix = sample_idx ? sample_idx[j] : j;
    const int32_t v = pix / w, u = pix % w;
    /* depth2pc: pc = (Kinv @ [u+.5, v+.5, 1]) * z ; mask on pc.z  (mapping_utils.py:239-249) */
    const double x2 = u + 0.5, y2 = v + 0.5;
    const double z = depth_at(depth, depth_u16, pix);
    const double px = dot3(kinv + 0, x2, y2, 1.0) * z;
    const double py = dot3(kinv + 3, x2, y2, 1.0) * z;
    const double pz = dot3(kinv + 6, x2, y2, 1.0) * z;
    if (!(pz > min_depth && pz < max_depth)) continue;
    /* transform_pc: pose @ [p; 1]  (mapping_utils.py:311-315) */
    const double gx = dot4h(tf + 0, px, py, pz);
    const double gy = dot4h(tf + 4, px, py, pz);
    const double gz = dot4h(tf + 8, px, py, pz);
    int64_t row, col, hh;
    int32_t pos_raw[3];
    int height_oob = 0;
    if (b->mode == 0) {
      /* base_pos2grid_id_3d (mapping_utils.py:345-349) */
      row = trunc_i64(half - (double)trunc_i64(gx / cs));
      col = trunc_i64(half - (double)trunc_i64(gy / cs));
      hh = trunc_i64(gz / cs);
      /* _out_of_range (vlmap_builder.py:283-284) */
      if (col >= gs || row >= gs || hh >= vh || col < 0 || row < 0 || hh < 0) continue;
      pos_raw[0] = (int32_t)row; pos_raw[1] = (int32_t)col; pos_raw[2] = (int32_t)hh;
    } else {
      /* row, height, col = np.round(((p - pcd_min) / cs)).astype(int)  (vlmap_builder_multi_floor.py:146) */
      row = round_i64((gx - b->origin[0]) / cs);
      hh = round_i64((gy - b->origin[1]) / cs);
      col = round_i64((gz - b->origin[2]) / cs);
      if (row >= b->n0 || col >= b->n1) continue; /* :151-153, the only test the reference makes */
      pos_raw[0] = (int32_t)row; pos_raw[1] = (int32_t)col; pos_raw[2] = (int32_t)hh; /* grid_pos keeps these (:176) */
      /* numpy indexing: negative indices wrap once, anything else raises IndexError */
      if (row < 0) row += b->n0;
      if (col < 0) col += b->n1;
      if (hh < 0) hh += b->n2;
      if (row < 0 || col < 0) { b->n_oob++; continue; } /* height_map[row, col] raises (:155) */
      height_oob = hh < 0 || hh >= b->n2;               /* occupied_ids[row, col, height] would raise (:175) */
    }
    /* project_point with the RGB calibration (vlmap_builder.py:141-142): no bounds check in the
     * reference (python negative indices wrap); we wrap the same way when rgb is given */
    uint8_t rgb_v[3] = {0, 0, 0};
    if (rgb) {
      const double q0 = dot3(k + 0, px, py, pz), q1 = dot3(k + 3, px, py, pz), q2 = dot3(k + 6, px, py, pz);
      int64_t rx = trunc_i64(q0 / q2 - 0.5), ry = trunc_i64(q1 / q2 - 0.5);
      if (rx < 0) rx += w;
      if (ry < 0) ry += h;
      if (rx >= 0 && rx < w && ry >= 0 && ry < h) memcpy(rgb_v, rgb + ((size_t)ry * w + rx) * 3, 3);
    }
    /* project_point with the feature camera (vlmap_builder.py:143) */
    const double f0 = dot3(kfeat + 0, px, py, pz), f1 = dot3(kfeat + 3, px, py, pz), f2 = dot3(kfeat + 6, px, py, pz);
    const int64_t fx = trunc_i64(f0 / f2 - 0.5), fy = trunc_i64(f1 / f2 - 0.5);
    /* alpha (vlmap_builder.py:156-158) */
    const double rsq = (px * px + py * py) + pz * pz;
    const double alpha = exp(-rsq / (2 * 0.6));
    if (fx < 0 || fy < 0 || fx >= fw || fy >= fh) continue; /* :161 */
    if (height_oob) { b->n_oob++; continue; }
    b->n_accepted++;
    const float* fp = feat + (size_t)fy * fw + fx; /* pix_feats[0, :, py, px], stride FH*FW */
    const size_t cell = ((size_t)row * b->n1 + col) * b->n2 + hh;
    int32_t id = b->occupied_ids[cell];
    if (id == -1) { /* :164-170 */
      if (b->max_id >= b->capacity) return -1;
      id = (int32_t)b->max_id;
      b->occupied_ids[cell] = id;
      float* g = b->grid_feat + (size_t)id * dim;
      for (int32_t c = 0; c < dim; ++c) g[c] = (float)((double)fp[c * plane] * alpha);
      memcpy(b->grid_rgb + (size_t)id * 3, rgb_v, 3);
      b->weight[id] = (float)((double)b->weight[id] + alpha);
      b->grid_pos[id * 3 + 0] = pos_raw[0];
      b->grid_pos[id * 3 + 1] = pos_raw[1];
      b->grid_pos[id * 3 + 2] = pos_raw[2];
      b->max_id++;
    } else { /* :171-178 */
      float* g = b->grid_feat + (size_t)id * dim;
      const float wf = b->weight[id];
      const double den = (double)wf + alpha;
      for (int32_t c = 0; c < dim; ++c) {
        const float t1 = g[c] * wf; /* float32 array * float32 scalar */
        g[c] = (float)(((double)t1 + (double)fp[c * plane] * alpha) / den);
      }
      uint8_t* cr = b->grid_rgb + (size_t)id * 3;
      for (int c = 0; c < 3; ++c) {
        const float t1 = (float)cr[c] * wf; /* uint8 array * float32 scalar -> float32 */
        const double val = ((double)t1 + (double)rgb_v[c] * alpha) / den;
        cr[c] = (uint8_t)val; /* C cast on store into the uint8 array */
      }
      b->weight[id] = (float)den;
    }
  }
  return 0;
}

int oracle_builder_add_frame(oracle_builder* b, const float* depth, int32_t h, int32_t w, const float* feat,
                             int32_t fh, int32_t fw, const uint8_t* rgb, const int32_t* sample_idx,
                             int32_t n_samples, const double* kinv, const double* k, const double* kfeat,
                             const double* tf, double min_depth, double max_depth) {
  return add_frame_impl(b, depth, 0, h, w, feat, fh, fw, rgb, sample_idx, n_samples, kinv, k, kfeat, tf, min_depth,
                        max_depth);
}
int oracle_builder_add_frame_u16(oracle_builder* b, const uint16_t* depth_mm, int32_t h, int32_t w, const float* feat,
                                 int32_t fh, int32_t fw, const uint8_t* rgb, const int32_t* sample_idx,
                                 int32_t n_samples, const double* kinv, const double* k, const double* kfeat,
                                 const double* tf, double min_depth, double max_depth) {
  return add_frame_impl(b, depth_mm, 1, h, w, feat, fh, fw, rgb, sample_idx, n_samples, kinv, k, kfeat, tf, min_depth,
                        max_depth);
}

/* First pass of create_global_map (vlmap_builder_multi_floor.py:97-118): component-wise min / max of the
 * transformed valid sampled points, merged into minmax[0..2] (min) / minmax[3..5] (max); returns the count. */
int64_t oracle_frame_bounds(const void* depth, int depth_u16, int32_t h, int32_t w, const int32_t* sample_idx,
                            int32_t n_samples, const double* kinv, const double* tf, double min_depth,
                            double max_depth, double* minmax) {
  int64_t cnt = 0;
  (void)h;
  for (int32_t j = 0; j < n_samples; ++j) {
    const int32_t pix = sample_idx ? sample_idx[j] : j;
    const int32_t v = pix / w, u = pix % w;
    const double x2 = u + 0.5, y2 = v + 0.5;
    const double z = depth_at(depth, depth_u16, pix);
    const double px = dot3(kinv + 0, x2, y2, 1.0) * z;
    const double py = dot3(kinv + 3, x2, y2, 1.0) * z;
    const double pz = dot3(kinv + 6, x2, y2, 1.0) * z;
    if (!(pz > min_depth && pz < max_depth)) continue;
    const double g[3] = {dot4h(tf + 0, px, py, pz), dot4h(tf + 4, px, py, pz), dot4h(tf + 8, px, py, pz)};
    for (int c = 0; c < 3; ++c) {
      if (g[c] < minmax[c]) minmax[c] = g[c];
      if (g[c] > minmax[3 + c]) minmax[3 + c] = g[c];
    }
    ++cnt;
  }
  return cnt;
}

/* ---------------------------------------------------------------------------------------------
 * Index path: canonical scores  s = fl32( sum_k (double)a_k (double)b_k ), k ascending
 * (restates `map_feats @ text_feats.T`, avlmaps/utils/clip_utils.py:229,240; see DESIGN.md for why the
 * accumulation is fp64), then the reference's argmax (avlmaps/map/vlmap.py:123).
 * ------------------------------------------------------------------------------------------- */
void oracle_scores(const float* feat, int64_t n, int32_t d, const float* q, int32_t nq, const float* scale,
                   int normalize, float* out) {
  for (int64_t i = 0; i < n; ++i) {
    const float* a = feat + i * d;
    float inv = 1.0f;
    if (normalize) {
      double s = 0.0;
      for (int32_t kk = 0; kk < d; ++kk) s += (double)a[kk] * (double)a[kk];
      const float nrm = (float)sqrt(s);
      inv = nrm > 0.0f ? 1.0f / nrm : 0.0f;
    }
    for (int32_t j = 0; j < nq; ++j) {
      const float* bq = q + (size_t)j * d;
      double acc = 0.0;
      for (int32_t kk = 0; kk < d; ++kk) acc += (double)a[kk] * (double)bq[kk];
      float s = (float)acc;
      if (normalize) s = s * inv;
      if (scale) s = s * scale[j];
      out[i * nq + j] = s;
    }
  }
}

void oracle_argmax(const float* scores, int64_t n, int32_t nq, int32_t* out) {
  for (int64_t i = 0; i < n; ++i) {
    const float* s = scores + i * nq;
    int32_t best = 0;
    for (int32_t j = 1; j < nq; ++j)
      if (s[j] > s[best]) best = j; /* first maximum, like np.argmax */
    out[i] = best;
  }
}
