// Skip-path resampling on channels-last data (libs/scale.py:7-45): FeaturePooling, bilinear x2,
// AvgPool 2x2.  Pure gathers, HBM-bound; every kernel is written output-stationary so the backward
// passes need no atomics and are deterministic.
#include "common.cuh"

// ---- FeaturePooling (scale.py:12-16) --------------------------------------------------------
// The reference views the NCHW-CONTIGUOUS memory of x as [B, Cout, H, W, r] and averages the last
// axis: out[b,o,h,w] = mean_k flat_b[o*r*HW + (h*W+w)*r + k].  flat index f <-> (channel f / HW,
// pixel f % HW), which we evaluate against the channels-last storage.
__global__ void k_featpool_fwd(const float* __restrict__ x, float* __restrict__ y, size_t n_out, int hw, int c_in, int c_out, int r) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const float inv = 1.0f / r;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) {
    const int o = (int)(i % c_out);
    const size_t bp = i / c_out;
    const int p = (int)(bp % hw);
    const size_t b = bp / hw;
    const size_t f0 = (size_t)o * r * hw + (size_t)p * r;
    float acc = 0.0f;
    for (int k = 0; k < r; ++k) {
      const size_t f = f0 + k;
      acc += x[(b * hw + f % hw) * c_in + f / hw];
    }
    y[i] = acc * inv;
  }
}
__global__ void k_featpool_bwd(const float* __restrict__ g, float* __restrict__ dx, size_t n_in, int hw, int c_in, int c_out, int r) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const float inv = 1.0f / r;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += stride) {
    const int c = (int)(i % c_in);
    const size_t bp = i / c_in;
    const int p = (int)(bp % hw);
    const size_t b = bp / hw;
    const size_t f = (size_t)c * hw + p;          // flat NCHW index inside the sample
    const size_t o = f / ((size_t)r * hw);
    const size_t q = (f % ((size_t)r * hw)) / r;  // output pixel
    dx[i] = g[(b * hw + q) * c_out + o] * inv;
  }
}
extern "C" int lb_featpool_fwd(const float* x, float* y, int batch, int h, int w, int c_in, int c_out, lb_stream_t s) {
  LB_REQUIRE(x && y && batch > 0 && h > 0 && w > 0 && c_out > 0 && c_in % c_out == 0);
  const size_t n = (size_t)batch * h * w * c_out;
  k_featpool_fwd<<<lb_grid_1d(n, 256), 256, 0, lb_s(s)>>>(x, y, n, h * w, c_in, c_out, c_in / c_out);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_featpool_bwd(const float* g, float* dx, int batch, int h, int w, int c_in, int c_out, lb_stream_t s) {
  LB_REQUIRE(g && dx && batch > 0 && h > 0 && w > 0 && c_out > 0 && c_in % c_out == 0);
  const size_t n = (size_t)batch * h * w * c_in;
  k_featpool_bwd<<<lb_grid_1d(n, 256), 256, 0, lb_s(s)>>>(g, dx, n, h * w, c_in, c_out, c_in / c_out);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ---- bilinear x2, align_corners=False (scale.py:37-38) --------------------------------------
// src = max(0, (dst + 0.5)/2 - 0.5); i0 = floor(src); i1 = min(i0+1, n-1); lambda = src - i0
__device__ __forceinline__ void lb_up2_src(int d, int n, int& i0, int& i1, float& lam) {
  float src = fmaxf(0.0f, (d + 0.5f) * 0.5f - 0.5f);
  i0 = (int)src;
  i1 = min(i0 + 1, n - 1);
  lam = src - i0;
}
// V = 4: one thread = 4 consecutive channels (128-bit accesses, 4x fewer index computations); V = 1: any layout
template <int V> struct VecT;
template <> struct VecT<1> { typedef float type; };
template <> struct VecT<4> { typedef float4 type; };
__device__ __forceinline__ float vmix(float a, float b, float c, float d, float w0, float w1, float w2, float w3) {
  return w0 * a + w1 * b + w2 * c + w3 * d;
}
__device__ __forceinline__ float4 vmix(float4 a, float4 b, float4 c, float4 d, float w0, float w1, float w2, float w3) {
  return make_float4(w0 * a.x + w1 * b.x + w2 * c.x + w3 * d.x, w0 * a.y + w1 * b.y + w2 * c.y + w3 * d.y,
                     w0 * a.z + w1 * b.z + w2 * c.z + w3 * d.z, w0 * a.w + w1 * b.w + w2 * c.w + w3 * d.w);
}
__device__ __forceinline__ void vfma(float& acc, float w, float v) { acc = fmaf(w, v, acc); }
__device__ __forceinline__ void vfma(float4& acc, float w, float4 v) {
  acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
}
__device__ __forceinline__ float vzero(float) { return 0.0f; }
__device__ __forceinline__ float4 vzero(float4) { return make_float4(0.f, 0.f, 0.f, 0.f); }

template <int V>
__global__ void __launch_bounds__(256) k_up2_fwd(const float* __restrict__ xs, float* __restrict__ ys, size_t n_out, int h, int w, int c) {
  typedef typename VecT<V>::type T;
  const T* __restrict__ x = reinterpret_cast<const T*>(xs);
  T* __restrict__ y = reinterpret_cast<T*>(ys);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int ow = 2 * w, oh = 2 * h;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) {
    const int ch = (int)(i % c);
    size_t t = i / c;
    const int ox = (int)(t % ow); t /= ow;
    const int oy = (int)(t % oh);
    const size_t b = t / oh;
    int y0, y1, x0, x1; float ly, lx;
    lb_up2_src(oy, h, y0, y1, ly);
    lb_up2_src(ox, w, x0, x1, lx);
    const T* xb = x + b * h * w * c + ch;
    y[i] = vmix(xb[((size_t)y0 * w + x0) * c], xb[((size_t)y0 * w + x1) * c], xb[((size_t)y1 * w + x0) * c],
                xb[((size_t)y1 * w + x1) * c], (1.0f - ly) * (1.0f - lx), (1.0f - ly) * lx, ly * (1.0f - lx), ly * lx);
  }
}
// weight with which source index m receives from destination index d along one axis
__device__ __forceinline__ float lb_up2_weight(int d, int n, int m) {
  int i0, i1; float lam;
  lb_up2_src(d, n, i0, i1, lam);
  return (i0 == m ? 1.0f - lam : 0.0f) + (i1 == m ? lam : 0.0f);
}
template <int V>
__global__ void __launch_bounds__(256) k_up2_bwd(const float* __restrict__ gs, float* __restrict__ dxs, size_t n_in, int h, int w, int c) {
  typedef typename VecT<V>::type T;
  const T* __restrict__ g = reinterpret_cast<const T*>(gs);
  T* __restrict__ dx = reinterpret_cast<T*>(dxs);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int ow = 2 * w, oh = 2 * h;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += stride) {
    const int ch = (int)(i % c);
    size_t t = i / c;
    const int ix = (int)(t % w); t /= w;
    const int iy = (int)(t % h);
    const size_t b = t / h;
    const T* gb = g + b * oh * ow * c + ch;
    T acc = vzero(T());
    for (int dy = max(0, 2 * iy - 2); dy <= min(oh - 1, 2 * iy + 2); ++dy) {
      const float wy = lb_up2_weight(dy, h, iy);
      if (wy == 0.0f) continue;
      for (int dxp = max(0, 2 * ix - 2); dxp <= min(ow - 1, 2 * ix + 2); ++dxp) {
        const float wx = lb_up2_weight(dxp, w, ix);
        if (wx != 0.0f) vfma(acc, wy * wx, gb[((size_t)dy * ow + dxp) * c]);
      }
    }
    dx[i] = acc;
  }
}
extern "C" int lb_upsample2x_fwd(const float* x, float* y, int batch, int h, int w, int c, lb_stream_t s) {
  LB_REQUIRE(x && y && batch > 0 && h > 0 && w > 0 && c > 0);
  const size_t n = (size_t)batch * h * w * c * 4;
  if ((c & 3) == 0 && lb_aligned16(x) && lb_aligned16(y))
    k_up2_fwd<4><<<lb_grid_1d(n / 4, 256), 256, 0, lb_s(s)>>>(x, y, n / 4, h, w, c / 4);
  else
    k_up2_fwd<1><<<lb_grid_1d(n, 256), 256, 0, lb_s(s)>>>(x, y, n, h, w, c);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_upsample2x_bwd(const float* g, float* dx, int batch, int h, int w, int c, lb_stream_t s) {
  LB_REQUIRE(g && dx && batch > 0 && h > 0 && w > 0 && c > 0);
  const size_t n = (size_t)batch * h * w * c;
  if ((c & 3) == 0 && lb_aligned16(g) && lb_aligned16(dx))
    k_up2_bwd<4><<<lb_grid_1d(n / 4, 256), 256, 0, lb_s(s)>>>(g, dx, n / 4, h, w, c / 4);
  else
    k_up2_bwd<1><<<lb_grid_1d(n, 256), 256, 0, lb_s(s)>>>(g, dx, n, h, w, c);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ---- AvgPool 2x2 / stride 2 (scale.py:40) ----------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256) k_avgpool2_fwd(const float* __restrict__ xs, float* __restrict__ ys, size_t n_out, int h, int w, int c) {
  typedef typename VecT<V>::type T;
  const T* __restrict__ x = reinterpret_cast<const T*>(xs);
  T* __restrict__ y = reinterpret_cast<T*>(ys);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int oh = h / 2, ow = w / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) {
    const int ch = (int)(i % c);
    size_t t = i / c;
    const int ox = (int)(t % ow); t /= ow;
    const int oy = (int)(t % oh);
    const size_t b = t / oh;
    const T* p = x + ((b * h + 2 * oy) * w + 2 * ox) * c + ch;
    y[i] = vmix(p[0], p[c], p[(size_t)w * c], p[(size_t)w * c + c], 0.25f, 0.25f, 0.25f, 0.25f);
  }
}
template <int V>
__global__ void __launch_bounds__(256) k_avgpool2_bwd(const float* __restrict__ gs, float* __restrict__ dxs, size_t n_in, int h, int w, int c) {
  typedef typename VecT<V>::type T;
  const T* __restrict__ g = reinterpret_cast<const T*>(gs);
  T* __restrict__ dx = reinterpret_cast<T*>(dxs);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int oh = h / 2, ow = w / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += stride) {
    const int ch = (int)(i % c);
    size_t t = i / c;
    const int ix = (int)(t % w); t /= w;
    const int iy = (int)(t % h);
    const size_t b = t / h;
    const int oy = iy >> 1, ox = ix >> 1;
    T acc = vzero(T());
    if (oy < oh && ox < ow) vfma(acc, 0.25f, g[((b * oh + oy) * ow + ox) * c + ch]);
    dx[i] = acc;
  }
}
extern "C" int lb_avgpool2_fwd(const float* x, float* y, int batch, int h, int w, int c, lb_stream_t s) {
  LB_REQUIRE(x && y && batch > 0 && h > 1 && w > 1 && c > 0);
  const size_t n = (size_t)batch * (h / 2) * (w / 2) * c;
  if ((c & 3) == 0 && lb_aligned16(x) && lb_aligned16(y))
    k_avgpool2_fwd<4><<<lb_grid_1d(n / 4, 256), 256, 0, lb_s(s)>>>(x, y, n / 4, h, w, c / 4);
  else
    k_avgpool2_fwd<1><<<lb_grid_1d(n, 256), 256, 0, lb_s(s)>>>(x, y, n, h, w, c);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_avgpool2_bwd(const float* g, float* dx, int batch, int h, int w, int c, lb_stream_t s) {
  LB_REQUIRE(g && dx && batch > 0 && h > 1 && w > 1 && c > 0);
  const size_t n = (size_t)batch * h * w * c;
  if ((c & 3) == 0 && lb_aligned16(g) && lb_aligned16(dx))
    k_avgpool2_bwd<4><<<lb_grid_1d(n / 4, 256), 256, 0, lb_s(s)>>>(g, dx, n / 4, h, w, c / 4);
  else
    k_avgpool2_bwd<1><<<lb_grid_1d(n, 256), 256, 0, lb_s(s)>>>(g, dx, n, h, w, c);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
