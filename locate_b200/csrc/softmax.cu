// Softmax kernels (libs/attention.py:35,47).
//  * over PIXELS for every (b,c) of a channels-last [B][P][C] tensor: SelfAttention's softmax over
//    HW (dim=-1 of [B,F,HW]).  Rows of the logical problem are strided by C in memory, so a CTA owns
//    (b, 32-channel slab): lanes run over channels (128 B coalesced) and warps stride over pixels.
//  * over the contiguous last axis of [rows][cols]: feature attention's Softmax(dim=1) on [B,F,1,1].
// fp32 throughout: the values are ~1/HW (6e-5 at 128^2) and feed (gamma*att+1), bf16 would erase them.
#include "common.cuh"

#define SM_WARPS 16

__global__ void __launch_bounds__(32 * SM_WARPS) k_softmax_pixels_fwd(const float* __restrict__ x, float* __restrict__ y,
                                                                      int pixels, int channels) {
  __shared__ float s_max[SM_WARPS][33];
  __shared__ float s_sum[SM_WARPS][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const bool ok = c < channels;
  const size_t base = (size_t)blockIdx.y * pixels * channels + (ok ? c : 0);
  float m = -INFINITY, l = 0.0f;
  if (ok) {
    for (int p = w; p < pixels; p += SM_WARPS) {
      const float v = x[base + (size_t)p * channels];
      const float nm = fmaxf(m, v);
      l = l * __expf(m - nm) + __expf(v - nm);
      m = nm;
    }
  }
  s_max[w][lane] = m;
  s_sum[w][lane] = l;
  __syncthreads();
  float gm = -INFINITY;
#pragma unroll
  for (int i = 0; i < SM_WARPS; ++i) gm = fmaxf(gm, s_max[i][lane]);
  float gl = 0.0f;
#pragma unroll
  for (int i = 0; i < SM_WARPS; ++i) {
    const float mi = s_max[i][lane];
    gl += (mi == -INFINITY) ? 0.0f : s_sum[i][lane] * __expf(mi - gm);
  }
  const float inv = 1.0f / gl;
  if (ok) {
    for (int p = w; p < pixels; p += SM_WARPS) {
      const size_t i = base + (size_t)p * channels;
      y[i] = __expf(x[i] - gm) * inv;
    }
  }
}

extern "C" int lb_softmax_pixels_fwd(const float* x, float* y, int batch, int pixels, int channels, lb_stream_t s) {
  LB_REQUIRE(x && y && batch > 0 && pixels > 0 && channels > 0 && batch <= 65535);
  k_softmax_pixels_fwd<<<dim3((channels + 31) / 32, batch), 32 * SM_WARPS, 0, lb_s(s)>>>(x, y, pixels, channels);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// dx = y * (g - sum_p y*g)
__global__ void __launch_bounds__(32 * SM_WARPS) k_softmax_pixels_bwd(const float* __restrict__ y, const float* __restrict__ g,
                                                                      float* __restrict__ dx, int pixels, int channels) {
  __shared__ float s_dot[SM_WARPS][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const bool ok = c < channels;
  const size_t base = (size_t)blockIdx.y * pixels * channels + (ok ? c : 0);
  float d = 0.0f;
  if (ok) {
    for (int p = w; p < pixels; p += SM_WARPS) {
      const size_t i = base + (size_t)p * channels;
      d = fmaf(y[i], g[i], d);
    }
  }
  s_dot[w][lane] = d;
  __syncthreads();
  float tot = 0.0f;
#pragma unroll
  for (int i = 0; i < SM_WARPS; ++i) tot += s_dot[i][lane];
  if (ok) {
    for (int p = w; p < pixels; p += SM_WARPS) {
      const size_t i = base + (size_t)p * channels;
      dx[i] = y[i] * (g[i] - tot);
    }
  }
}
extern "C" int lb_softmax_pixels_bwd(const float* y, const float* g, float* dx, int batch, int pixels, int channels, lb_stream_t s) {
  LB_REQUIRE(y && g && dx && batch > 0 && pixels > 0 && channels > 0 && batch <= 65535);
  k_softmax_pixels_bwd<<<dim3((channels + 31) / 32, batch), 32 * SM_WARPS, 0, lb_s(s)>>>(y, g, dx, pixels, channels);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// contiguous rows: one warp per row
__global__ void __launch_bounds__(128) k_softmax_rows_fwd(const float* __restrict__ x, float* __restrict__ y, int rows, int cols) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * cols;
  float m = -INFINITY;
  for (int j = lane; j < cols; j += 32) m = fmaxf(m, xr[j]);
  m = lb_warp_max(m);
  float l = 0.0f;
  for (int j = lane; j < cols; j += 32) l += __expf(xr[j] - m);
  l = lb_warp_sum(l);
  const float inv = 1.0f / l;
  for (int j = lane; j < cols; j += 32) y[(size_t)row * cols + j] = __expf(xr[j] - m) * inv;
}
__global__ void __launch_bounds__(128) k_softmax_rows_bwd(const float* __restrict__ y, const float* __restrict__ g,
                                                         float* __restrict__ dx, int rows, int cols) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const size_t o = (size_t)row * cols;
  float d = 0.0f;
  for (int j = lane; j < cols; j += 32) d = fmaf(y[o + j], g[o + j], d);
  d = lb_warp_sum(d);
  for (int j = lane; j < cols; j += 32) dx[o + j] = y[o + j] * (g[o + j] - d);
}
extern "C" int lb_softmax_rows_fwd(const float* x, float* y, int rows, int cols, lb_stream_t s) {
  LB_REQUIRE(x && y && rows > 0 && cols > 0);
  k_softmax_rows_fwd<<<(rows + 3) / 4, 128, 0, lb_s(s)>>>(x, y, rows, cols);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_softmax_rows_bwd(const float* y, const float* g, float* dx, int rows, int cols, lb_stream_t s) {
  LB_REQUIRE(y && g && dx && rows > 0 && cols > 0);
  k_softmax_rows_bwd<<<(rows + 3) / 4, 128, 0, lb_s(s)>>>(y, g, dx, rows, cols);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
