// Spectral norm (libs/spectral_norm.py:21-32): one power iteration per forward call, and the weight
// gradient epilogue that folds d(1/sigma)/dW back in.  Two passes over W per forward, HBM-bound:
//   pass 1  t = W^T u  (column sums weighted by u)     -> v = t / (|t| + eps)
//   pass 2  s = W v    (row dot products)              -> u = s / (|s| + eps), sigma = u.s = |s|^2/(|s|+eps)
// The GEMMs read W_bar untouched and apply 1/sigma in their epilogue (conv(x, W/sigma) == conv(x, W)/sigma).
#include "common.cuh"

#define SN_EPS 1e-12f

// Deterministic by construction (the reference's torch.mv is): every row split writes its own partial column sums and
// the normalise kernel adds the splits in a fixed order -- no floating-point atomics anywhere in the iteration.
// tpart[split][j] = sum_{i in row split} W[i][j] * u[i];   grid = (col tiles of 256, row splits)
__global__ void __launch_bounds__(256) k_sn_wt_u(const float* __restrict__ w, const float* __restrict__ u, float* __restrict__ tpart,
                                                int height, int width, int rows_per_split) {
  lb_pdl_enter();
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= width) return;
  const int i0 = blockIdx.y * rows_per_split, i1 = min(height, i0 + rows_per_split);
  float acc = 0.0f;
  for (int i = i0; i < i1; ++i) acc = fmaf(w[(size_t)i * width + j], __ldg(u + i), acc);
  tpart[(size_t)blockIdx.y * width + j] = acc;
}

// src[i] <- sum of its `nsplit` partials (stride n, fixed order); dst = src / (|src| + eps); also writes sigma if non-null.  One CTA.
__global__ void __launch_bounds__(1024) k_sn_normalize(float* __restrict__ src, float* __restrict__ dst, int n, int nsplit,
                                                      float* __restrict__ sigma_out) {
  lb_pdl_enter();
  __shared__ float scratch[32];
  __shared__ float s_norm;
  float acc = 0.0f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float v = src[i];
    for (int sp = 1; sp < nsplit; ++sp) v += src[(size_t)sp * n + i];
    src[i] = v;
    acc = fmaf(v, v, acc);
  }
  acc = lb_block_sum(acc, scratch);
  if (threadIdx.x == 0) s_norm = sqrtf(acc);
  __syncthreads();
  const float nrm = s_norm;
  const float inv = 1.0f / (nrm + SN_EPS);
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i] * inv;
  if (sigma_out && threadIdx.x == 0) {
    const float sigma = nrm * nrm * inv;                 // u.(W v) with u = s/(|s|+eps)
    sigma_out[0] = sigma;
    sigma_out[1] = 1.0f / sigma;
  }
}

// s[i] = sum_j W[i][j] v[j]; one warp per row (rows are contiguous)
__global__ void __launch_bounds__(256) k_sn_w_v(const float* __restrict__ w, const float* __restrict__ v, float* __restrict__ sv,
                                               int height, int width) {
  lb_pdl_enter();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= height) return;
  const float* wr = w + (size_t)row * width;
  float acc = 0.0f;
  for (int j = lane; j < width; j += 32) acc = fmaf(wr[j], __ldg(v + j), acc);
  acc = lb_warp_sum(acc);
  if (lane == 0) sv[row] = acc;
}

static void sn_splits(int height, int width, int& splits, int& rps) {
  const int col_tiles = (width + 255) / 256;
  splits = (LB_SMS * 2 + col_tiles - 1) / col_tiles;
  if (splits > 16) splits = 16;              // bounds the partial buffer; 16 x col_tiles CTAs still cover the GPU for any wide W
  if (splits > height) splits = height;
  if (splits < 1) splits = 1;
  rps = (height + splits - 1) / splits;
  splits = (height + rps - 1) / rps;
}
extern "C" size_t lb_sn_power_iter_work_floats(int height, int width) {
  if (height <= 0 || width <= 0) return 0;
  int splits, rps;
  sn_splits(height, width, splits, rps);
  return (size_t)splits * width + height;
}
extern "C" int lb_sn_power_iter(const float* w, int height, int width, float* u, float* v, float* sigma_out, float* work,
                                lb_stream_t s) {
  LB_REQUIRE(w && u && v && sigma_out && work && height > 0 && width > 0);
  int splits, rps;
  sn_splits(height, width, splits, rps);
  float* t = work;                              // [splits][width]
  float* sv = work + (size_t)splits * width;    // [height]
  const int col_tiles = (width + 255) / 256;
  lb_launch(k_sn_wt_u, dim3(col_tiles, splits), 256, 0, lb_s(s), w, u, t, height, width, rps);
  LB_LAUNCH_CHECK();
  lb_launch(k_sn_normalize, 1, 1024, 0, lb_s(s), t, v, width, splits, nullptr);
  LB_LAUNCH_CHECK();
  lb_launch(k_sn_w_v, (height + 7) / 8, 256, 0, lb_s(s), w, v, sv, height, width);
  LB_LAUNCH_CHECK();
  lb_launch(k_sn_normalize, 1, 1024, 0, lb_s(s), sv, u, height, 1, sigma_out);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ---- weight gradient epilogue --------------------------------------------------------------
// dwn is either in the master layout (packed_taps = 0) or tap-major packed [taps][d0][d1] as written by
// lb_wgrad_tc (master [d0][d1][taps]); the master index k maps to the packed offset below.
// Index decoding uses multiply-shift division by (width, taps): a 64-bit divide per element made these two kernels
// instruction-bound (they only stream W-sized arrays).
struct SnIdx { LbFastDiv d_width, d_taps; int width, height, taps, d1; };
__device__ __forceinline__ void sn_decode(const SnIdx& x, int k, int& i, int& j, size_t& off) {
  lb_fast_divmod(x.d_width, k, i, j);
  if (x.taps == 0) { off = (size_t)k; return; }
  int j1, tap;
  lb_fast_divmod(x.d_taps, j, j1, tap);
  off = ((size_t)tap * x.height + i) * x.d1 + j1;
}
__global__ void __launch_bounds__(256) k_sn_dot(const float* __restrict__ a, const float* __restrict__ b, int n, const SnIdx x,
                                               double* __restrict__ out, double* __restrict__ stat_work) {
  lb_pdl_enter();
  __shared__ double scratch[32];
  const int stride = gridDim.x * blockDim.x;
  float part = 0.0f;
  double acc = 0.0;
  int cnt = 0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    int i, j; size_t off;
    sn_decode(x, k, i, j, off);
    part = fmaf(a[off], b[k], part);
    if (++cnt == 16) { acc += (double)part; part = 0.0f; cnt = 0; }     // short fp32 runs, fp64 across them
  }
  acc += (double)part;
  acc = lb_block_sum(acc, scratch);
  lb_grid_sum2_ordered(acc, 0.0, stat_work, out, scratch);             // fixed-order grid sum: reproducible gradients
}
// grad[i][j] += dwn[i][j]/sigma - dot/sigma^2 * u[i] v[j].  When the layer's u / v are trainable (see lb_sn_weight_grad):
// du[i] += c * s_fwd[i] and cacc += c with c = dL/dsigma = -dot/sigma^2.
__global__ void __launch_bounds__(256) k_sn_wgrad(const float* __restrict__ dwn, const float* __restrict__ u, const float* __restrict__ v,
                                                 const float* __restrict__ sigma, const double* __restrict__ dot,
                                                 float* __restrict__ grad, int n, const SnIdx x, const float* __restrict__ s_fwd,
                                                 float* __restrict__ du, float* __restrict__ cacc) {
  lb_pdl_enter();
  const float inv = __ldg(sigma + 1);
  const float coef = (float)(dot[0] * (double)inv * (double)inv);
  const int stride = gridDim.x * blockDim.x;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    int i, j; size_t off;
    sn_decode(x, k, i, j, off);
    grad[k] += fmaf(dwn[off], inv, -coef * __ldg(u + i) * __ldg(v + j));
    if (du && k < x.height) du[k] = fmaf(-coef, __ldg(s_fwd + k), du[k]);
    if (cacc && k == 0) cacc[0] -= coef;
  }
}
// contiguous (packed_taps = 0) fast paths: 4 consecutive elements of one matrix row per thread
__global__ void __launch_bounds__(256) k_sn_dot4(const float4* __restrict__ a, const float4* __restrict__ b, int n4,
                                                double* __restrict__ out, double* __restrict__ stat_work) {
  lb_pdl_enter();
  __shared__ double scratch[32];
  const int stride = gridDim.x * blockDim.x;
  double acc = 0.0;
  float part = 0.0f;
  int cnt = 0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += stride) {
    const float4 x = __ldcs(a + k), y = __ldg(b + k);
    part = fmaf(x.x, y.x, fmaf(x.y, y.y, fmaf(x.z, y.z, fmaf(x.w, y.w, part))));
    if (++cnt == 4) { acc += (double)part; part = 0.0f; cnt = 0; }
  }
  acc += (double)part;
  acc = lb_block_sum(acc, scratch);
  lb_grid_sum2_ordered(acc, 0.0, stat_work, out, scratch);
}
__global__ void __launch_bounds__(256) k_sn_wgrad4(const float4* __restrict__ dwn, const float* __restrict__ u, const float* __restrict__ v,
                                                  const float* __restrict__ sigma, const double* __restrict__ dot,
                                                  float4* __restrict__ grad, int n4, int height, LbFastDiv d_w4,
                                                  const float* __restrict__ s_fwd, float* __restrict__ du, float* __restrict__ cacc) {
  lb_pdl_enter();
  const float inv = __ldg(sigma + 1);
  const float coef = (float)(dot[0] * (double)inv * (double)inv);
  const int stride = gridDim.x * blockDim.x;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += stride) {
    int i, j4;
    lb_fast_divmod(d_w4, k, i, j4);
    const float cu = -coef * __ldg(u + i);
    const float4 vv = __ldg(reinterpret_cast<const float4*>(v) + j4);
    const float4 d = __ldcs(dwn + k);
    float4 g = grad[k];
    g.x += fmaf(d.x, inv, cu * vv.x);
    g.y += fmaf(d.y, inv, cu * vv.y);
    g.z += fmaf(d.z, inv, cu * vv.z);
    g.w += fmaf(d.w, inv, cu * vv.w);
    grad[k] = g;
    if (du && k < height) du[k] = fmaf(-coef, __ldg(s_fwd + k), du[k]);
    if (cacc && k == 0) cacc[0] -= coef;
  }
}
// w == NULL: dot_out[0] already holds sum dwn * W (lb_wgrad_tc computes it while it writes dwn).
extern "C" int lb_sn_weight_grad(const float* dwn, const float* w, const float* u, const float* v, const float* sigma,
                                 float* grad, int height, int width, int packed_taps, double* dot_out, double* stat_work,
                                 const float* s_fwd, float* du, float* cacc, lb_stream_t s) {
  LB_REQUIRE(dwn && u && v && sigma && grad && dot_out && stat_work && height > 0 && width > 0 && packed_taps >= 0);
  LB_REQUIRE(w || packed_taps == 0);
  LB_REQUIRE(packed_taps == 0 || width % packed_taps == 0);
  LB_REQUIRE((s_fwd && du && cacc) || (!s_fwd && !du && !cacc));
  const size_t n = (size_t)height * width;
  LB_REQUIRE(n < ((size_t)1 << 31) - ((size_t)1 << 24));
  SnIdx x;
  x.width = width; x.height = height; x.taps = packed_taps; x.d1 = packed_taps ? width / packed_taps : width;
  x.d_width = lb_make_fastdiv(width); x.d_taps = lb_make_fastdiv(packed_taps ? packed_taps : 1);
  const bool vec = packed_taps == 0 && width % 4 == 0 && height <= (int)(n / 4) &&
                   !((reinterpret_cast<uintptr_t>(dwn) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(v) |
                      (w ? reinterpret_cast<uintptr_t>(w) : 0)) & 15);
  if (w) {
    if (vec)
      lb_launch(k_sn_dot4, lb_grid_1d(n / 4, 256, 4), 256, 0, lb_s(s), reinterpret_cast<const float4*>(dwn), reinterpret_cast<const float4*>(w),
                                                                  (int)(n / 4), dot_out, stat_work);
    else
      lb_launch(k_sn_dot, lb_grid_1d(n, 256, 2), 256, 0, lb_s(s), dwn, w, (int)n, x, dot_out, stat_work);
    LB_LAUNCH_CHECK();
  }
  if (vec)
    lb_launch(k_sn_wgrad4, lb_grid_1d(n / 4, 256), 256, 0, lb_s(s), reinterpret_cast<const float4*>(dwn), u, v, sigma, dot_out,
                                                              reinterpret_cast<float4*>(grad), (int)(n / 4), height,
                                                              lb_make_fastdiv(width / 4), s_fwd, du, cacc);
  else
    lb_launch(k_sn_wgrad, lb_grid_1d(n, 256), 256, 0, lb_s(s), dwn, u, v, sigma, dot_out, grad, (int)n, x, s_fwd, du, cacc);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ---- batched power iteration: every spectral-normed layer of a model in 4 launches ------------------------
// The per-layer version costs 4 tiny launches per layer per forward (~270 per discriminator pass).  u/v/sigma of a
// layer depend only on (W, u), so all layers can be iterated up front.  Work is flattened into items so that a
// 37 M-element weight and a 96-element weight share one grid without load imbalance.
struct LbSnLayerDev {
  const float* w; float* u; float* v; float* dv;
  int height, width, t_off, s_off, nsplit, rows_per_split;
};
// phase 1 item: layer, first column, first row, row count.  Row split r0 / rows_per_split of the layer writes its own
// partial tpart[split][j] = sum_rows W[i][j] u[i] (plain stores: deterministic, nothing to zero first)
__global__ void __launch_bounds__(256) k_snb_wt_u(const LbSnLayerDev* __restrict__ layers, const int4* __restrict__ items,
                                                 float* __restrict__ scratch) {
  lb_pdl_enter();
  const int4 it = items[blockIdx.x];
  const LbSnLayerDev L = layers[it.x];
  const int j = it.y + 4 * threadIdx.x;                    // 4 consecutive columns per thread
  if (j >= L.width) return;
  const int i1 = min(L.height, it.z + it.w);
  float* dst = scratch + (size_t)L.t_off + (size_t)(it.z / L.rows_per_split) * L.width + j;
  if ((L.width & 3) == 0 && (reinterpret_cast<uintptr_t>(L.w) & 15) == 0) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int i = it.z; i < i1; ++i) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(L.w + (size_t)i * L.width + j));
      const float ui = __ldg(L.u + i);
      acc.x = fmaf(wv.x, ui, acc.x); acc.y = fmaf(wv.y, ui, acc.y); acc.z = fmaf(wv.z, ui, acc.z); acc.w = fmaf(wv.w, ui, acc.w);
    }
    dst[0] = acc.x; dst[1] = acc.y; dst[2] = acc.z; dst[3] = acc.w;
  } else {
    const int cols = min(4, L.width - j);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = it.z; i < i1; ++i) {
      const float ui = __ldg(L.u + i);
      for (int e = 0; e < cols; ++e) acc[e] = fmaf(L.w[(size_t)i * L.width + j + e], ui, acc[e]);
    }
    for (int e = 0; e < cols; ++e) dst[e] = acc[e];
  }
}
// per layer: dst = src/(|src|+eps); phase 2 (v from t) and phase 4 (u from s, sigma)
__global__ void __launch_bounds__(1024) k_snb_normalize(const LbSnLayerDev* __restrict__ layers, float* __restrict__ scratch,
                                                      float* __restrict__ s_out, int phase, float* __restrict__ sigma_out) {
  lb_pdl_enter();
  __shared__ float red[32];
  __shared__ float s_norm;
  const LbSnLayerDev L = layers[blockIdx.x];
  float* src = phase == 2 ? scratch + L.t_off : s_out + L.s_off;
  float* dst = phase == 2 ? L.v : L.u;
  const int n = phase == 2 ? L.width : L.height;
  const int nsplit = phase == 2 ? L.nsplit : 1;
  float acc = 0.0f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float v = src[i];
    int sp = 1;
    for (; sp + 4 <= nsplit; sp += 4) {                                   // fixed order over the row splits, loads in flight together
      const float a = src[(size_t)sp * n + i], b = src[(size_t)(sp + 1) * n + i], c = src[(size_t)(sp + 2) * n + i],
                  d = src[(size_t)(sp + 3) * n + i];
      v = (((v + a) + b) + c) + d;
    }
    for (; sp < nsplit; ++sp) v += src[(size_t)sp * n + i];
    src[i] = v;
    acc = fmaf(v, v, acc);
  }
  acc = lb_block_sum(acc, red);
  if (threadIdx.x == 0) s_norm = sqrtf(acc);
  __syncthreads();
  const float nrm = s_norm;
  const float inv = 1.0f / (nrm + SN_EPS);
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i] * inv;
  if (phase == 4 && threadIdx.x == 0) {
    const float sigma = nrm * nrm * inv;
    sigma_out[2 * blockIdx.x] = sigma;
    sigma_out[2 * blockIdx.x + 1] = 1.0f / sigma;
  }
}
// phase 3 item: (layer, first row); 8 rows per CTA, one warp per row.  s_out[s_off + i] = sum_j W[i][j] v[j]
__global__ void __launch_bounds__(256) k_snb_w_v(const LbSnLayerDev* __restrict__ layers, const int2* __restrict__ items,
                                                float* __restrict__ scratch) {
  lb_pdl_enter();
  const int2 it = items[blockIdx.x];
  const LbSnLayerDev L = layers[it.x];
  const int row = it.y + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= L.height) return;
  const float* wr = L.w + (size_t)row * L.width;
  float acc = 0.0f;
  if ((L.width & 3) == 0 && ((reinterpret_cast<uintptr_t>(L.w) | reinterpret_cast<uintptr_t>(L.v)) & 15) == 0) {
#pragma unroll 2
    for (int j = 4 * lane; j < L.width; j += 128) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + j));
      const float4 vv = __ldg(reinterpret_cast<const float4*>(L.v + j));
      acc = fmaf(wv.x, vv.x, fmaf(wv.y, vv.y, fmaf(wv.z, vv.z, fmaf(wv.w, vv.w, acc))));
    }
  } else {
    for (int j = lane; j < L.width; j += 32) acc = fmaf(wr[j], __ldg(L.v + j), acc);
  }
  acc = lb_warp_sum(acc);
  if (lane == 0) scratch[L.s_off + row] = acc;        // `scratch` is s_out here
}
extern "C" int lb_sn_power_iter_batched(const void* layers_dev, int n_layers, const void* items1_dev, int n_items1,
                                        const void* items3_dev, int n_items3, float* scratch, float* s_out,
                                        float* sigma_out, lb_stream_t s) {
  LB_REQUIRE(layers_dev && items1_dev && items3_dev && scratch && s_out && sigma_out && n_layers > 0 && n_items1 > 0 && n_items3 > 0);
  // every partial slot is written before it is read: no zero fill
  const LbSnLayerDev* layers = reinterpret_cast<const LbSnLayerDev*>(layers_dev);
  lb_launch(k_snb_wt_u, n_items1, 256, 0, lb_s(s), layers, reinterpret_cast<const int4*>(items1_dev), scratch);
  LB_LAUNCH_CHECK();
  lb_launch(k_snb_normalize, n_layers, 1024, 0, lb_s(s), layers, scratch, s_out, 2, sigma_out);
  LB_LAUNCH_CHECK();
  lb_launch(k_snb_w_v, n_items3, 256, 0, lb_s(s), layers, reinterpret_cast<const int2*>(items3_dev), s_out);
  LB_LAUNCH_CHECK();
  lb_launch(k_snb_normalize, n_layers, 1024, 0, lb_s(s), layers, scratch, s_out, 4, sigma_out);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// Trainable u / v (the reference's main.py:172 quirk: `dis.requires_grad_(True)` also switches on the discriminator's
// weight_u / weight_v, whose sigma = u.(W v) then hands them gradients):  dv += C * W^T u with the LIVE u, where
// C = sum over the backward passes of dL/dsigma (accumulated in cacc by lb_sn_weight_grad; reset here).  One extra pass
// over the weights per optimizer step, reusing the pass-1 kernel of the power iteration.
__global__ void __launch_bounds__(512) k_snb_dv(const LbSnLayerDev* __restrict__ layers, const float* __restrict__ scratch,
                                               float* __restrict__ cacc) {
  lb_pdl_enter();
  const LbSnLayerDev L = layers[blockIdx.x];
  const float c = cacc[blockIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) cacc[blockIdx.x] = 0.0f;
  if (!L.dv || c == 0.0f) return;
  const float* src = scratch + L.t_off;
  for (int j = threadIdx.x; j < L.width; j += blockDim.x) {
    float t = src[j];
    for (int sp = 1; sp < L.nsplit; ++sp) t += src[(size_t)sp * L.width + j];
    L.dv[j] = fmaf(c, t, L.dv[j]);
  }
}
extern "C" int lb_sn_uv_grad_batched(const void* layers_dev, int n_layers, const void* items1_dev, int n_items1, float* scratch,
                                     float* cacc, lb_stream_t s) {
  LB_REQUIRE(layers_dev && items1_dev && scratch && cacc && n_layers > 0 && n_items1 > 0);
  const LbSnLayerDev* layers = reinterpret_cast<const LbSnLayerDev*>(layers_dev);
  lb_launch(k_snb_wt_u, n_items1, 256, 0, lb_s(s), layers, reinterpret_cast<const int4*>(items1_dev), scratch);
  LB_LAUNCH_CHECK();
  lb_launch(k_snb_dv, n_layers, 512, 0, lb_s(s), layers, scratch, cacc);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
