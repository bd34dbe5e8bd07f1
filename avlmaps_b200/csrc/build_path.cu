// Map-build path: back-projection + alpha-weighted feature fusion, B200-native.
//
// Replaces the per-frame body of VLMapBuilder.create_mobile_base_map
// (reference avlmaps/map/vlmap_builder.py:129-178) and the helpers it calls per point
// (depth2pc mapping_utils.py:226-251, transform_pc :305-315, base_pos2grid_id_3d :345-349,
// project_point :599-605).  The reference loop is sequential and order dependent (first touch of a
// cell stores feat*alpha with weight alpha, later touches average); the order-free form used here
// (SURVEY.md section 0.4, appendix A):
//     key(point)   = (frame_seq << 32) | position in the frame's sample list
//     first(cell)  = min key over the points of the cell                     -> atomicMin
//     voxel id     = rank of first(cell) among all cells = running count of "winner" points in
//                    (frame, sample) order                                   -> ordered scan per frame
//     grid_feat    = (alpha_first^2 f_first + sum_{others} alpha f) / sum alpha
//     weight       = sum alpha
// Kernels per frame: geometry (fp64, no FMA contraction) -> winner count -> scan -> id assignment ->
// scatter-reduce (warp per point, float4 vector reds into the voxel row).
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "avl_internal.h"

namespace avl {
namespace {

struct FrameGeom {
  double kinv[9], k[9], kfeat[9], tf[16];
  double min_depth, max_depth, cs, half_gs;
  int32_t h, w, fh, fw, gs, vh;
  int32_t has_rgb;
};

constexpr int kScanBlock = 1024;
constexpr unsigned long long kNoKey = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ double dot3(const double* m, double x, double y, double z) {
  // (m0*x + m1*y) + m2*z, every operation rounded separately like numpy's float64 matmul here
  return __dadd_rn(__dadd_rn(__dmul_rn(m[0], x), __dmul_rn(m[1], y)), __dmul_rn(m[2], z));
}
__device__ __forceinline__ long long trunc_ll(double v) {
  // python int(): toward zero; far-out values are clamped (the range tests reject them anyway)
  if (!(v > -9.0e15)) return -(1ll << 60);
  if (!(v < 9.0e15)) return (1ll << 60);
  return __double2ll_rz(v);
}

// ---------------------------------------------------------------- geometry + first-touch keys
__global__ void __launch_bounds__(256)
geom_kernel(const FrameGeom g, const float* __restrict__ depth, const int32_t* __restrict__ sample_idx,
            int32_t n_samples, uint32_t frame_seq, unsigned long long* __restrict__ first_key,
            int32_t* __restrict__ s_cell, int32_t* __restrict__ s_fpix, float* __restrict__ s_alpha,
            int32_t* __restrict__ s_rgbpix) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_samples; j += gridDim.x * blockDim.x) {
    const int pix = sample_idx ? sample_idx[j] : j;
    const int v = pix / g.w, u = pix - v * g.w;
    const double x2 = u + 0.5, y2 = v + 0.5;
    const double z = static_cast<double>(depth[pix]);
    // depth2pc: pc = (Kinv @ [u+.5, v+.5, 1]) * z   (mapping_utils.py:239-246)
    const double px = __dmul_rn(dot3(g.kinv + 0, x2, y2, 1.0), z);
    const double py = __dmul_rn(dot3(g.kinv + 3, x2, y2, 1.0), z);
    const double pz = __dmul_rn(dot3(g.kinv + 6, x2, y2, 1.0), z);
    int cell = -1, fpix = 0, rgbpix = -1;
    float alpha = 0.f;
    if (pz > g.min_depth && pz < g.max_depth) {  // mapping_utils.py:247-249
      // transform_pc: pose @ [p; 1]   (mapping_utils.py:311-315)
      const double gx = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(g.tf[0], px), __dmul_rn(g.tf[1], py)), __dmul_rn(g.tf[2], pz)), g.tf[3]);
      const double gy = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(g.tf[4], px), __dmul_rn(g.tf[5], py)), __dmul_rn(g.tf[6], pz)), g.tf[7]);
      const double gz = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(g.tf[8], px), __dmul_rn(g.tf[9], py)), __dmul_rn(g.tf[10], pz)), g.tf[11]);
      // base_pos2grid_id_3d (mapping_utils.py:345-349): double truncation toward zero
      const long long row = trunc_ll(__dsub_rn(g.half_gs, static_cast<double>(trunc_ll(__ddiv_rn(gx, g.cs)))));
      const long long col = trunc_ll(__dsub_rn(g.half_gs, static_cast<double>(trunc_ll(__ddiv_rn(gy, g.cs)))));
      const long long hh = trunc_ll(__ddiv_rn(gz, g.cs));
      if (!(col >= g.gs || row >= g.gs || hh >= g.vh || col < 0 || row < 0 || hh < 0)) {  // vlmap_builder.py:283
        // project_point with the feature camera (vlmap_builder.py:143, mapping_utils.py:599-605)
        const double f2 = dot3(g.kfeat + 6, px, py, pz);
        const long long fx = trunc_ll(__dsub_rn(__ddiv_rn(dot3(g.kfeat + 0, px, py, pz), f2), 0.5));
        const long long fy = trunc_ll(__dsub_rn(__ddiv_rn(dot3(g.kfeat + 3, px, py, pz), f2), 0.5));
        if (!(fx < 0 || fy < 0 || fx >= g.fw || fy >= g.fh)) {  // vlmap_builder.py:161
          cell = static_cast<int>((row * g.gs + col) * g.vh + hh);
          fpix = static_cast<int>(fy * g.fw + fx);
          // alpha = exp(-||p||^2 / (2 * 0.6))   (vlmap_builder.py:156-158)
          const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py)), __dmul_rn(pz, pz));
          alpha = static_cast<float>(exp(__ddiv_rn(-rsq, 1.2)));
          if (g.has_rgb) {  // vlmap_builder.py:141-142: no bounds check; negative indices wrap like numpy
            const double q2 = dot3(g.k + 6, px, py, pz);
            long long rx = trunc_ll(__dsub_rn(__ddiv_rn(dot3(g.k + 0, px, py, pz), q2), 0.5));
            long long ry = trunc_ll(__dsub_rn(__ddiv_rn(dot3(g.k + 3, px, py, pz), q2), 0.5));
            if (rx < 0) rx += g.w;
            if (ry < 0) ry += g.h;
            if (rx >= 0 && rx < g.w && ry >= 0 && ry < g.h) rgbpix = static_cast<int>(ry * g.w + rx);
          }
          const unsigned long long key = (static_cast<unsigned long long>(frame_seq) << 32) | static_cast<uint32_t>(j);
          atomicMin(first_key + cell, key);
        }
      }
    }
    s_cell[j] = cell;
    s_fpix[j] = fpix;
    s_alpha[j] = alpha;
    s_rgbpix[j] = rgbpix;
  }
}

__device__ __forceinline__ bool is_winner(const unsigned long long* first_key, int cell, uint32_t frame_seq, int j) {
  return cell >= 0 &&
         first_key[cell] == ((static_cast<unsigned long long>(frame_seq) << 32) | static_cast<uint32_t>(j));
}

// ---------------------------------------------------------------- ordered id assignment
__global__ void __launch_bounds__(kScanBlock)
winner_count_kernel(const int32_t* __restrict__ s_cell, int32_t n_samples, uint32_t frame_seq,
                    const unsigned long long* __restrict__ first_key, uint32_t* __restrict__ block_cnt,
                    unsigned long long* __restrict__ n_accepted) {
  const int j = blockIdx.x * kScanBlock + threadIdx.x;
  const int cell = j < n_samples ? s_cell[j] : -1;
  const int wins = __syncthreads_count(is_winner(first_key, cell, frame_seq, j));
  const int acc = __syncthreads_count(cell >= 0);
  if (threadIdx.x == 0) {
    block_cnt[blockIdx.x] = static_cast<uint32_t>(wins);
    if (acc) atomicAdd(n_accepted, static_cast<unsigned long long>(acc));
  }
}

// one block: exclusive scan of block_cnt, offset by max_id; max_id += total
__global__ void __launch_bounds__(kScanBlock)
winner_scan_kernel(uint32_t* __restrict__ block_cnt, int32_t nblocks, unsigned long long* __restrict__ max_id) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = static_cast<uint32_t>(*max_id);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nblocks; base += kScanBlock) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < nblocks ? block_cnt[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;  // inclusive totals of the warps
    }
    __syncthreads();
    const uint32_t before = carry + (warp ? warp_tot[warp - 1] : 0u) + (x - v);
    if (i < nblocks) block_cnt[i] = before;  // now: first id of the block
    __syncthreads();
    if (threadIdx.x == kScanBlock - 1) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *max_id = carry;
}

__global__ void __launch_bounds__(kScanBlock)
assign_ids_kernel(const int32_t* __restrict__ s_cell, int32_t n_samples, uint32_t frame_seq,
                  const unsigned long long* __restrict__ first_key, const uint32_t* __restrict__ block_base,
                  int32_t gs, int32_t vh, int64_t capacity, int32_t* __restrict__ occupied_ids,
                  int32_t* __restrict__ grid_pos) {
  __shared__ uint32_t warp_tot[32];
  const int j = blockIdx.x * kScanBlock + threadIdx.x;
  const int cell = j < n_samples ? s_cell[j] : -1;
  const bool win = is_winner(first_key, cell, frame_seq, j);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t b = __ballot_sync(0xffffffffu, win);
  if (lane == 0) warp_tot[warp] = __popc(b);
  __syncthreads();
  if (warp == 0) {
    uint32_t t = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    warp_tot[lane] = t;
  }
  __syncthreads();
  if (win) {
    const int64_t id = static_cast<int64_t>(block_base[blockIdx.x]) + (warp ? warp_tot[warp - 1] : 0u) +
                       __popc(b & ((1u << lane) - 1u));
    if (id < capacity) {
      occupied_ids[cell] = static_cast<int32_t>(id);  // vlmap_builder.py:165
      const int hh = cell % vh, rc = cell / vh;
      grid_pos[id * 3 + 0] = rc / gs;                  // vlmap_builder.py:169
      grid_pos[id * 3 + 1] = rc % gs;
      grid_pos[id * 3 + 2] = hh;
    }
  }
}

// ---------------------------------------------------------------- feature layout
// (D, P) -> (P, D), P = FH*FW pixels.  64 pixels x 64 channels per block through shared memory:
// 256-byte contiguous reads per channel row, 256-byte contiguous writes per pixel row (float4 both
// ways when P and D allow it).
__global__ void __launch_bounds__(256)
chw_to_hwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int32_t d, int64_t p) {
  __shared__ float tile[64][65];
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int c0 = blockIdx.y * 64;
  const int t = threadIdx.x;
  const bool vec_in = (p & 3) == 0, vec_out = (d & 3) == 0;
  // load: row = channel, 16 threads x float4 cover 64 pixels
  {
    const int q4 = (t & 15) * 4, r0 = t >> 4;
#pragma unroll
    for (int rr = 0; rr < 64; rr += 16) {
      const int c = c0 + r0 + rr;
      const int64_t pp = p0 + q4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < d) {
        const float* s = src + static_cast<int64_t>(c) * p + pp;
        if (vec_in && pp + 3 < p) v = __ldg(reinterpret_cast<const float4*>(s));
        else {
          if (pp < p) v.x = s[0];
          if (pp + 1 < p) v.y = s[1];
          if (pp + 2 < p) v.z = s[2];
          if (pp + 3 < p) v.w = s[3];
        }
      }
      tile[r0 + rr][q4] = v.x; tile[r0 + rr][q4 + 1] = v.y; tile[r0 + rr][q4 + 2] = v.z; tile[r0 + rr][q4 + 3] = v.w;
    }
  }
  __syncthreads();
  // store: row = pixel, 16 threads x float4 cover 64 channels
  {
    const int c4 = (t & 15) * 4, r0 = t >> 4;
#pragma unroll
    for (int rr = 0; rr < 64; rr += 16) {
      const int64_t pp = p0 + r0 + rr;
      const int c = c0 + c4;
      if (pp >= p) continue;
      const float4 v = make_float4(tile[c4][r0 + rr], tile[c4 + 1][r0 + rr], tile[c4 + 2][r0 + rr], tile[c4 + 3][r0 + rr]);
      float* o = dst + pp * d + c;
      if (vec_out && c + 3 < d) *reinterpret_cast<float4*>(o) = v;
      else {
        if (c < d) o[0] = v.x;
        if (c + 1 < d) o[1] = v.y;
        if (c + 2 < d) o[2] = v.z;
        if (c + 3 < d) o[3] = v.w;
      }
    }
  }
}

// ---------------------------------------------------------------- scatter-reduce
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// One warp per accepted point: num[id, :] += w * feat[fpix, :], den[id] += alpha, rgb likewise.
// w = alpha^2 for the point that first touched the cell (vlmap_builder.py:164-170 stores feat*alpha
// with weight alpha), alpha otherwise (:171-178).
__global__ void __launch_bounds__(256)
scatter_kernel(const float* __restrict__ feat_hwc, int32_t d, const uint8_t* __restrict__ rgb,
               const int32_t* __restrict__ s_cell, const int32_t* __restrict__ s_fpix,
               const float* __restrict__ s_alpha, const int32_t* __restrict__ s_rgbpix, int32_t n_samples,
               uint32_t frame_seq, const unsigned long long* __restrict__ first_key,
               const int32_t* __restrict__ occupied_ids, int64_t capacity, float* __restrict__ num,
               float* __restrict__ den, float* __restrict__ rgb_acc) {
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n_samples; j += nwarps) {
    const int cell = s_cell[j];
    if (cell < 0) continue;
    const int64_t id = occupied_ids[cell];
    if (id < 0 || id >= capacity) continue;
    const float alpha = s_alpha[j];
    const bool win = is_winner(first_key, cell, frame_seq, j);
    const float wgt = win ? alpha * alpha : alpha;
    const float* f = feat_hwc + static_cast<int64_t>(s_fpix[j]) * d;
    float* o = num + id * d;
    if ((d & 3) == 0) {
      const float4* f4 = reinterpret_cast<const float4*>(f);
      for (int c = lane; c < (d >> 2); c += 32) {
        const float4 v = __ldg(f4 + c);
        red_add_v4(o + 4 * c, v.x * wgt, v.y * wgt, v.z * wgt, v.w * wgt);
      }
    } else {
      for (int c = lane; c < d; c += 32) atomicAdd(o + c, f[c] * wgt);
    }
    if (lane == 0) atomicAdd(den + id, alpha);
    if (rgb && lane >= 1 && lane <= 3) {
      const int rp = s_rgbpix[j];
      const float cv = rp >= 0 ? static_cast<float>(rgb[static_cast<int64_t>(rp) * 3 + (lane - 1)]) : 0.f;
      atomicAdd(rgb_acc + id * 3 + (lane - 1), cv * alpha);
    }
  }
}

// ---------------------------------------------------------------- export
__global__ void __launch_bounds__(256)
export_feat_kernel(const float* __restrict__ num, const float* __restrict__ den, int64_t v, int32_t d,
                   float* __restrict__ out) {
  const int64_t total = v * d;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = __fdiv_rn(num[i], den[i / d]);
}
__global__ void __launch_bounds__(256)
export_rgb_kernel(const float* __restrict__ rgb_acc, const float* __restrict__ den, int64_t v,
                  uint8_t* __restrict__ out) {
  const int64_t total = v * 3;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float x = __fdiv_rn(rgb_acc[i], den[i / 3]);
    out[i] = static_cast<uint8_t>(fminf(fmaxf(x, 0.f), 255.f));  // truncation, like the uint8 store
  }
}
__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = v;
}
__global__ void fill_u64_kernel(unsigned long long* p, int64_t n, unsigned long long v) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = v;
}

}  // namespace
}  // namespace avl

using namespace avl;

struct avl_builder {
  avl_grid_spec spec;
  int num_sms = 148;
  int64_t cells = 0;
  int64_t capacity = 0;
  uint32_t frame_seq = 0;
  int64_t id_upper = 0;  // host-side upper bound of max_id (avoids a sync per frame)
  // persistent device state
  unsigned long long* first_key = nullptr;
  int32_t* occupied_ids = nullptr;
  float* num = nullptr;
  float* den = nullptr;
  float* rgb_acc = nullptr;
  int32_t* grid_pos = nullptr;
  unsigned long long* counters = nullptr;  // [0] max_id, [1] n_accepted
  // per-frame scratch (grown on demand)
  int32_t *s_cell = nullptr, *s_fpix = nullptr, *s_rgbpix = nullptr;
  float* s_alpha = nullptr;
  uint32_t* block_cnt = nullptr;
  int64_t scratch_samples = 0;
  // staging of host inputs
  float* d_depth = nullptr; size_t depth_elems = 0;
  float* d_feat = nullptr; size_t feat_elems = 0;
  float* d_feat_t = nullptr; size_t feat_t_elems = 0;
  uint8_t* d_rgb = nullptr; size_t rgb_bytes = 0;
  int32_t* d_sidx = nullptr; size_t sidx_elems = 0;
};

namespace {

template <typename T>
int grow(T** p, size_t* have, size_t need) {
  if (*have >= need) return AVL_OK;
  cudaFree(*p);
  *p = nullptr;
  *have = 0;
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(p), need * sizeof(T)));
  *have = need;
  return AVL_OK;
}

int alloc_rows(avl_builder* b, int64_t cap, float** num, float** den, float** rgb, int32_t** pos, cudaStream_t s) {
  const int d = b->spec.dim;
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(num), static_cast<size_t>(cap) * d * sizeof(float)));
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(den), static_cast<size_t>(cap) * sizeof(float)));
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(rgb), static_cast<size_t>(cap) * 3 * sizeof(float)));
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(pos), static_cast<size_t>(cap) * 3 * sizeof(int32_t)));
  AVL_CUDA(cudaMemsetAsync(*num, 0, static_cast<size_t>(cap) * d * sizeof(float), s));
  AVL_CUDA(cudaMemsetAsync(*den, 0, static_cast<size_t>(cap) * sizeof(float), s));
  AVL_CUDA(cudaMemsetAsync(*rgb, 0, static_cast<size_t>(cap) * 3 * sizeof(float), s));
  AVL_CUDA(cudaMemsetAsync(*pos, 0, static_cast<size_t>(cap) * 3 * sizeof(int32_t), s));
  return AVL_OK;
}

// capacity doubling, the analogue of _reserve_map_space (vlmap_builder.py:286-311)
int ensure_capacity(avl_builder* b, int64_t incoming, cudaStream_t s) {
  if (b->capacity >= b->cells) return AVL_OK;  // one row per cell already: cannot overflow
  if (b->id_upper + incoming <= b->capacity) {
    b->id_upper += incoming;
    return AVL_OK;
  }
  unsigned long long max_id = 0;
  AVL_CUDA(cudaMemcpyAsync(&max_id, b->counters, sizeof(max_id), cudaMemcpyDeviceToHost, s));
  AVL_CUDA(cudaStreamSynchronize(s));
  b->id_upper = static_cast<int64_t>(max_id);
  if (b->id_upper + incoming > b->capacity) {
    int64_t cap = b->capacity;
    while (cap < b->id_upper + incoming) cap *= 2;
    if (cap > b->cells) cap = std::max(b->cells, b->id_upper + incoming);
    float *num, *den, *rgb;
    int32_t* pos;
    int rc = alloc_rows(b, cap, &num, &den, &rgb, &pos, s);
    if (rc) return rc;
    const size_t v = static_cast<size_t>(b->id_upper), d = b->spec.dim;
    AVL_CUDA(cudaMemcpyAsync(num, b->num, v * d * sizeof(float), cudaMemcpyDeviceToDevice, s));
    AVL_CUDA(cudaMemcpyAsync(den, b->den, v * sizeof(float), cudaMemcpyDeviceToDevice, s));
    AVL_CUDA(cudaMemcpyAsync(rgb, b->rgb_acc, v * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    AVL_CUDA(cudaMemcpyAsync(pos, b->grid_pos, v * 3 * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    AVL_CUDA(cudaStreamSynchronize(s));
    cudaFree(b->num); cudaFree(b->den); cudaFree(b->rgb_acc); cudaFree(b->grid_pos);
    b->num = num; b->den = den; b->rgb_acc = rgb; b->grid_pos = pos;
    b->capacity = cap;
  }
  b->id_upper += incoming;
  return AVL_OK;
}

}  // namespace

extern "C" {

int avl_builder_create(const avl_grid_spec* spec, avl_builder** out) {
  AVL_ARG(spec != nullptr && out != nullptr, "NULL argument");
  *out = nullptr;
  AVL_ARG(spec->gs >= 1 && spec->vh >= 1 && spec->dim >= 1 && spec->cs > 0.0, "invalid grid spec");
  const int64_t cells = static_cast<int64_t>(spec->gs) * spec->gs * spec->vh;
  AVL_ARG(cells < (int64_t(1) << 31), "grid has more than 2^31 cells");
  int dev = 0, major = 0;
  AVL_CUDA(cudaGetDevice(&dev));
  AVL_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    set_error("avlmaps_b200 needs an sm_100a (B200) device");
    return AVL_ERR_UNSUPPORTED;
  }
  avl_builder* b = new avl_builder();
  b->spec = *spec;
  cudaDeviceGetAttribute(&b->num_sms, cudaDevAttrMultiProcessorCount, dev);
  b->cells = cells;
  b->capacity = spec->capacity > 0 ? std::min<int64_t>(spec->capacity, cells)
                                   : std::min<int64_t>(static_cast<int64_t>(spec->gs) * spec->gs, cells);
  int rc = AVL_OK;
  do {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&b->first_key), cells * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&b->occupied_ids), cells * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&b->counters), 2 * sizeof(unsigned long long));
    if (e != cudaSuccess) { rc = cuda_fail(e, "builder state", __FILE__, __LINE__); break; }
    fill_u64_kernel<<<1184, 256>>>(b->first_key, cells, kNoKey);
    fill_i32_kernel<<<1184, 256>>>(b->occupied_ids, cells, -1);  // vlmap_builder.py:204
    e = cudaMemset(b->counters, 0, 2 * sizeof(unsigned long long));
    if (e != cudaSuccess) { rc = cuda_fail(e, "builder init", __FILE__, __LINE__); break; }
    if ((rc = alloc_rows(b, b->capacity, &b->num, &b->den, &b->rgb_acc, &b->grid_pos, nullptr))) break;
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { rc = cuda_fail(e, "builder init", __FILE__, __LINE__); break; }
  } while (0);
  if (rc) {
    avl_builder_destroy(b);
    return rc;
  }
  *out = b;
  return AVL_OK;
}

int avl_builder_destroy(avl_builder* b) {
  if (!b) return AVL_OK;
  cudaFree(b->first_key); cudaFree(b->occupied_ids); cudaFree(b->num); cudaFree(b->den); cudaFree(b->rgb_acc);
  cudaFree(b->grid_pos); cudaFree(b->counters); cudaFree(b->s_cell); cudaFree(b->s_fpix); cudaFree(b->s_rgbpix);
  cudaFree(b->s_alpha); cudaFree(b->block_cnt); cudaFree(b->d_depth); cudaFree(b->d_feat); cudaFree(b->d_feat_t);
  cudaFree(b->d_rgb); cudaFree(b->d_sidx);
  delete b;
  return AVL_OK;
}

int avl_builder_add_frame(avl_builder* b, const avl_frame* f, int flags, void* stream) {
  AVL_ARG(b != nullptr && f != nullptr, "NULL argument");
  AVL_ARG(f->depth != nullptr && f->feat != nullptr, "depth / feat is NULL");
  AVL_ARG(f->h >= 1 && f->w >= 1 && f->fh >= 1 && f->fw >= 1, "invalid frame shape");
  AVL_ARG(static_cast<int64_t>(f->h) * f->w < (int64_t(1) << 31), "frame too large");
  AVL_ARG(f->feat_layout == AVL_FEAT_CHW || f->feat_layout == AVL_FEAT_HWC, "unknown feat_layout");
  const int64_t npix = static_cast<int64_t>(f->h) * f->w;
  const int32_t n_samples = f->sample_idx ? f->n_samples : static_cast<int32_t>(npix);
  AVL_ARG(n_samples >= 0 && n_samples <= npix, "n_samples out of range");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int d = b->spec.dim;
  const size_t fpix = static_cast<size_t>(f->fh) * f->fw;
  int rc;

  // ---- inputs on the device
  const float* depth = f->depth;
  const float* feat = f->feat;
  const uint8_t* rgb = f->rgb;
  const int32_t* sidx = f->sample_idx;
  if (!(flags & AVL_ON_DEVICE)) {
    if ((rc = grow(&b->d_depth, &b->depth_elems, static_cast<size_t>(npix)))) return rc;
    if ((rc = grow(&b->d_feat, &b->feat_elems, fpix * d))) return rc;
    AVL_CUDA(cudaMemcpyAsync(b->d_depth, f->depth, npix * sizeof(float), cudaMemcpyHostToDevice, s));
    AVL_CUDA(cudaMemcpyAsync(b->d_feat, f->feat, fpix * d * sizeof(float), cudaMemcpyHostToDevice, s));
    depth = b->d_depth;
    feat = b->d_feat;
    if (f->rgb) {
      if ((rc = grow(&b->d_rgb, &b->rgb_bytes, static_cast<size_t>(npix) * 3))) return rc;
      AVL_CUDA(cudaMemcpyAsync(b->d_rgb, f->rgb, npix * 3, cudaMemcpyHostToDevice, s));
      rgb = b->d_rgb;
    }
    if (f->sample_idx) {
      if ((rc = grow(&b->d_sidx, &b->sidx_elems, static_cast<size_t>(std::max(n_samples, 1))))) return rc;
      AVL_CUDA(cudaMemcpyAsync(b->d_sidx, f->sample_idx, static_cast<size_t>(n_samples) * sizeof(int32_t),
                               cudaMemcpyHostToDevice, s));
      sidx = b->d_sidx;
    }
  }
  if (n_samples == 0) {
    b->frame_seq++;
    return AVL_OK;
  }
  if (f->feat_layout == AVL_FEAT_CHW) {  // (1, D, FH, FW) -> pixel-major rows for the coalesced gather
    if ((rc = grow(&b->d_feat_t, &b->feat_t_elems, fpix * d))) return rc;
    dim3 grid(static_cast<unsigned>((fpix + 63) / 64), static_cast<unsigned>((d + 63) / 64));
    chw_to_hwc_kernel<<<grid, 256, 0, s>>>(feat, b->d_feat_t, d, static_cast<int64_t>(fpix));
    AVL_CUDA(cudaGetLastError());
    feat = b->d_feat_t;
  }

  // ---- scratch
  if (b->scratch_samples < n_samples) {
    cudaFree(b->s_cell); cudaFree(b->s_fpix); cudaFree(b->s_rgbpix); cudaFree(b->s_alpha); cudaFree(b->block_cnt);
    b->scratch_samples = 0;
    const size_t n = static_cast<size_t>(n_samples);
    AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->s_cell), n * sizeof(int32_t)));
    AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->s_fpix), n * sizeof(int32_t)));
    AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->s_rgbpix), n * sizeof(int32_t)));
    AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->s_alpha), n * sizeof(float)));
    AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->block_cnt), ((n + kScanBlock - 1) / kScanBlock) * sizeof(uint32_t)));
    b->scratch_samples = n_samples;
  }
  if ((rc = ensure_capacity(b, n_samples, s))) return rc;

  FrameGeom g;
  memcpy(g.kinv, f->kinv, sizeof(g.kinv));
  memcpy(g.k, f->k, sizeof(g.k));
  memcpy(g.kfeat, f->kfeat, sizeof(g.kfeat));
  memcpy(g.tf, f->tf, sizeof(g.tf));
  g.min_depth = f->min_depth;
  g.max_depth = f->max_depth;
  g.cs = b->spec.cs;
  g.half_gs = b->spec.gs / 2.0;
  g.h = f->h; g.w = f->w; g.fh = f->fh; g.fw = f->fw; g.gs = b->spec.gs; g.vh = b->spec.vh;
  g.has_rgb = rgb != nullptr;

  const int nblocks = (n_samples + kScanBlock - 1) / kScanBlock;
  const int geom_blocks = std::min((n_samples + 255) / 256, b->num_sms * 8);
  geom_kernel<<<geom_blocks, 256, 0, s>>>(g, depth, sidx, n_samples, b->frame_seq, b->first_key, b->s_cell,
                                          b->s_fpix, b->s_alpha, b->s_rgbpix);
  winner_count_kernel<<<nblocks, kScanBlock, 0, s>>>(b->s_cell, n_samples, b->frame_seq, b->first_key,
                                                     b->block_cnt, b->counters + 1);
  winner_scan_kernel<<<1, kScanBlock, 0, s>>>(b->block_cnt, nblocks, b->counters);
  assign_ids_kernel<<<nblocks, kScanBlock, 0, s>>>(b->s_cell, n_samples, b->frame_seq, b->first_key, b->block_cnt,
                                                   b->spec.gs, b->spec.vh, b->capacity, b->occupied_ids,
                                                   b->grid_pos);
  const int scatter_blocks = std::min((n_samples + 7) / 8, b->num_sms * 8);
  scatter_kernel<<<scatter_blocks, 256, 0, s>>>(feat, d, rgb, b->s_cell, b->s_fpix, b->s_alpha, b->s_rgbpix,
                                                n_samples, b->frame_seq, b->first_key, b->occupied_ids,
                                                b->capacity, b->num, b->den, b->rgb_acc);
  AVL_CUDA(cudaGetLastError());
  b->frame_seq++;
  if (!(flags & AVL_ON_DEVICE)) AVL_CUDA(cudaStreamSynchronize(s));  // staging buffers are reused per frame
  return AVL_OK;
}

static int read_counter(avl_builder* b, int which, int64_t* n, void* stream) {
  AVL_ARG(b != nullptr && n != nullptr, "NULL argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned long long v = 0;
  AVL_CUDA(cudaMemcpyAsync(&v, b->counters + which, sizeof(v), cudaMemcpyDeviceToHost, s));
  AVL_CUDA(cudaStreamSynchronize(s));
  *n = static_cast<int64_t>(v);
  return AVL_OK;
}
int avl_builder_num_voxels(avl_builder* b, int64_t* n, void* stream) { return read_counter(b, 0, n, stream); }
int avl_builder_num_accepted(avl_builder* b, int64_t* n, void* stream) { return read_counter(b, 1, n, stream); }

int avl_builder_export(avl_builder* b, float* grid_feat, int32_t* grid_pos, float* weight, int32_t* occupied_ids,
                       uint8_t* grid_rgb, int flags, void* stream) {
  AVL_ARG(b != nullptr, "builder is NULL");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int64_t v = 0;
  int rc = read_counter(b, 0, &v, stream);
  if (rc) return rc;
  const int d = b->spec.dim;
  const cudaMemcpyKind kind = (flags & AVL_ON_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (grid_feat && v > 0) {
    float* dst = grid_feat;
    if (!(flags & AVL_ON_DEVICE)) AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&dst), static_cast<size_t>(v) * d * sizeof(float)));
    export_feat_kernel<<<b->num_sms * 8, 256, 0, s>>>(b->num, b->den, v, d, dst);
    if (!(flags & AVL_ON_DEVICE)) {
      cudaError_t e = cudaMemcpyAsync(grid_feat, dst, static_cast<size_t>(v) * d * sizeof(float), cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      cudaFree(dst);
      if (e != cudaSuccess) return cuda_fail(e, "export grid_feat", __FILE__, __LINE__);
    }
  }
  if (grid_rgb && v > 0) {
    uint8_t* dst = grid_rgb;
    if (!(flags & AVL_ON_DEVICE)) AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&dst), static_cast<size_t>(v) * 3));
    export_rgb_kernel<<<b->num_sms * 2, 256, 0, s>>>(b->rgb_acc, b->den, v, dst);
    if (!(flags & AVL_ON_DEVICE)) {
      cudaError_t e = cudaMemcpyAsync(grid_rgb, dst, static_cast<size_t>(v) * 3, cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      cudaFree(dst);
      if (e != cudaSuccess) return cuda_fail(e, "export grid_rgb", __FILE__, __LINE__);
    }
  }
  if (grid_pos && v > 0) AVL_CUDA(cudaMemcpyAsync(grid_pos, b->grid_pos, static_cast<size_t>(v) * 3 * sizeof(int32_t), kind, s));
  if (weight && v > 0) AVL_CUDA(cudaMemcpyAsync(weight, b->den, static_cast<size_t>(v) * sizeof(float), kind, s));
  if (occupied_ids) AVL_CUDA(cudaMemcpyAsync(occupied_ids, b->occupied_ids, static_cast<size_t>(b->cells) * sizeof(int32_t), kind, s));
  AVL_CUDA(cudaGetLastError());
  if (!(flags & AVL_ON_DEVICE)) AVL_CUDA(cudaStreamSynchronize(s));
  return AVL_OK;
}

int avl_builder_to_map(avl_builder* b, void* stream, avl_map** out) {
  AVL_ARG(b != nullptr && out != nullptr, "NULL argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int64_t v = 0;
  int rc = read_counter(b, 0, &v, stream);
  if (rc) return rc;
  float* tmp = nullptr;
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&tmp), static_cast<size_t>(std::max<int64_t>(v, 1)) * b->spec.dim * sizeof(float)));
  if (v > 0) export_feat_kernel<<<b->num_sms * 8, 256, 0, s>>>(b->num, b->den, v, b->spec.dim, tmp);
  rc = avl_map_create(tmp, v, b->spec.dim, AVL_ON_DEVICE, stream, out);
  cudaFree(tmp);
  return rc;
}

}  // extern "C"
