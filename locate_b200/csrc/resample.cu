// Skip-path resampling on channels-last data (libs/scale.py:7-45): FeaturePooling, bilinear x2,
// AvgPool 2x2.  Pure gathers, HBM-bound; every kernel is written output-stationary so the backward
// passes need no atomics and are deterministic.
#include "common.cuh"

// ---- FeaturePooling (scale.py:12-16) --------------------------------------------------------
// The reference views the NCHW-CONTIGUOUS memory of x as [B, Cout, H, W, r] and averages the last
// axis: out[b,o,h,w] = mean_k flat_b[o*r*HW + (h*W+w)*r + k].  flat index f <-> (channel f / HW,
// pixel f % HW), which we evaluate against the channels-last storage.
// With m = HW / r (r divides HW on the reference's path): the r averaged elements of out[b,o,p] are channel
// ci = o*r + p / m at the consecutive pixels (p % m)*r + k, and dx[b,p,c] = g[b, (c % r)*m + p / r, c / r] / r.
// All index decoding is 32-bit multiply-shift division (the 64-bit divides of the first version made these gathers
// instruction-bound at a quarter of the HBM rate).
struct FeatPoolIdx { LbFastDiv d_c, d_hw, d_m, d_r; int hw, c_in, c_out, r, m; };
// forward: one thread = one INPUT channel ci of one pixel group pm (reads: consecutive threads on consecutive channels of
// r consecutive pixels, full lines); it produces out[b, (ci % r)*m + pm, ci / r].  d_c = c_in here.
__device__ __forceinline__ void lb_st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void lb_st2(lb_bf16* p, float a, float b) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b); }
template <typename T>
__global__ void __launch_bounds__(256) k_featpool_fwd(const T* __restrict__ x, T* __restrict__ y, int n_out, const FeatPoolIdx q) {
  lb_pdl_enter();
  const int stride = gridDim.x * blockDim.x;
  const float inv = 1.0f / q.r;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) {
    int t, ci, b, pm, o, pj;
    lb_fast_divmod(q.d_c, i, t, ci);
    lb_fast_divmod(q.d_m, t, b, pm);
    lb_fast_divmod(q.d_r, ci, o, pj);
    const T* src = x + ((size_t)b * q.hw + (size_t)pm * q.r) * q.c_in + ci;
    float acc = 0.0f;
    for (int k = 0; k < q.r; ++k) acc += lb_ld1(src + (size_t)k * q.c_in);
    lb_st1(y + ((size_t)b * q.hw + (size_t)pj * q.m + pm) * q.c_out + o, acc * inv);
  }
}
// r == 2 (every use on the reference's path: C -> C/2) with 4 input channels per thread: two 16-byte loads, two 8-byte
// stores (channels ci..ci+3 are outputs (o, o+1) of pixel groups pj = 0 and pj = 1).  d_c = c_in / 4 here.
template <typename T>
__global__ void __launch_bounds__(256) k_featpool_fwd_r2v4(const T* __restrict__ x, T* __restrict__ y, int n_items, const FeatPoolIdx q) {
  lb_pdl_enter();
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += stride) {
    int t, c4, b, pm;
    lb_fast_divmod(q.d_c, i, t, c4);
    lb_fast_divmod(q.d_m, t, b, pm);
    const T* src = x + ((size_t)b * q.hw + (size_t)pm * 2) * q.c_in + 4 * c4;
    const float4 a = lb_ld4(src), c = lb_ld4(src + q.c_in);
    const int o = 2 * c4;                                   // ci = 4*c4 + {0,1,2,3} -> (o, pj) = (2*c4, 0), (2*c4, 1), (2*c4+1, 0), (2*c4+1, 1)
    lb_st2(y + ((size_t)b * q.hw + pm) * q.c_out + o, 0.5f * (a.x + c.x), 0.5f * (a.z + c.z));              // pj = 0
    lb_st2(y + ((size_t)b * q.hw + q.m + pm) * q.c_out + o, 0.5f * (a.y + c.y), 0.5f * (a.w + c.w));        // pj = 1
  }
}
template <typename T>
__global__ void __launch_bounds__(256) k_featpool_bwd(const T* __restrict__ g, T* __restrict__ dx, int n_in, const FeatPoolIdx q) {
  lb_pdl_enter();
  const int stride = gridDim.x * blockDim.x;
  const float inv = 1.0f / q.r;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += stride) {
    int bp, c, b, p, o, cr, pq, pr;
    lb_fast_divmod(q.d_c, i, bp, c);              // d_c = c_in
    lb_fast_divmod(q.d_hw, bp, b, p);
    lb_fast_divmod(q.d_r, c, o, cr);
    lb_fast_divmod(q.d_r, p, pq, pr);
    lb_st1(dx + i, lb_ld1(g + ((size_t)b * q.hw + (size_t)cr * q.m + pq) * q.c_out + o) * inv);
  }
}
// r == 2 with 8 input channels per thread: channels c..c+7 of pixel p take outputs o = c/2 .. c/2+3 of pixel group p/2 --
// the even ones from output pixel p/2, the odd ones from output pixel m + p/2 -- so two 4-element loads and one 8-element
// store replace eight 2-byte gathers (the scalar kernel ran at 1 TB/s).  d_c = c_in / 8 here.
template <typename T>
__global__ void __launch_bounds__(256) k_featpool_bwd_r2v8(const T* __restrict__ g, T* __restrict__ dx, int n_items, const FeatPoolIdx q) {
  lb_pdl_enter();
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += stride) {
    int bp, c8, b, p;
    lb_fast_divmod(q.d_c, i, bp, c8);
    lb_fast_divmod(q.d_hw, bp, b, p);
    const T* src = g + ((size_t)b * q.hw + (p >> 1)) * q.c_out + 4 * c8;
    const float4 e = lb_ld4(src), o = lb_ld4(src + (size_t)q.m * q.c_out);
    T* dst = dx + (size_t)i * 8;
    lb_st4(dst, make_float4(0.5f * e.x, 0.5f * o.x, 0.5f * e.y, 0.5f * o.y));
    lb_st4(dst + 4, make_float4(0.5f * e.z, 0.5f * o.z, 0.5f * e.w, 0.5f * o.w));
  }
}
// generic fallbacks (r does not divide HW, or more than 2^31 items): 64-bit index arithmetic, any shape
template <typename T>
__global__ void k_featpool_fwd_generic(const T* __restrict__ x, T* __restrict__ y, size_t n_out, int hw, int c_in, int c_out, int r) {
  lb_pdl_enter();
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
      acc += lb_ld1(x + (b * hw + f % hw) * c_in + f / hw);
    }
    lb_st1(y + i, acc * inv);
  }
}
template <typename T>
__global__ void k_featpool_bwd_generic(const T* __restrict__ g, T* __restrict__ dx, size_t n_in, int hw, int c_in, int c_out, int r) {
  lb_pdl_enter();
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
    lb_st1(dx + i, lb_ld1(g + (b * hw + q) * c_out + o) * inv);
  }
}
static int featpool_idx(FeatPoolIdx* q, int batch, int h, int w, int c_in, int c_out, int c_fast, size_t n) {
  const int hw = h * w, r = c_in / c_out;
  if (hw % r != 0 || n >= ((size_t)1 << 31) - ((size_t)1 << 24)) return LB_EUNSUPPORTED;
  q->hw = hw; q->c_in = c_in; q->c_out = c_out; q->r = r; q->m = hw / r;
  q->d_c = lb_make_fastdiv(c_fast); q->d_hw = lb_make_fastdiv(hw); q->d_m = lb_make_fastdiv(q->m); q->d_r = lb_make_fastdiv(r);
  (void)batch;
  return LB_OK;
}
template <typename T>
static int featpool_fwd_t(const T* x, T* y, int batch, int h, int w, int c_in, int c_out, lb_stream_t s) {
  const size_t n = (size_t)batch * h * w * c_out;
  FeatPoolIdx q;
  if (featpool_idx(&q, batch, h, w, c_in, c_out, c_in, n) == LB_OK) {
    if (q.r == 2 && !(c_in & 3) && lb_vec4_ok(x) && !(reinterpret_cast<uintptr_t>(y) & (2 * sizeof(T) - 1))) {
      q.d_c = lb_make_fastdiv(c_in / 4);
      lb_launch(k_featpool_fwd_r2v4<T>, lb_grid_1d(n / 4, 256), 256, 0, lb_s(s), x, y, (int)(n / 4), q);     // n = B*m*c_in outputs, 4 per item
    } else {
      lb_launch(k_featpool_fwd<T>, lb_grid_1d(n, 256), 256, 0, lb_s(s), x, y, (int)n, q);
    }
  }
  else
    lb_launch(k_featpool_fwd_generic<T>, lb_grid_1d(n, 256), 256, 0, lb_s(s), x, y, n, h * w, c_in, c_out, c_in / c_out);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
template <typename T>
static int featpool_bwd_t(const T* g, T* dx, int batch, int h, int w, int c_in, int c_out, lb_stream_t s) {
  const size_t n = (size_t)batch * h * w * c_in;
  FeatPoolIdx q;
  if (featpool_idx(&q, batch, h, w, c_in, c_out, c_in, n) == LB_OK) {
    if (q.r == 2 && !(c_in & 7) && lb_vec4_ok(g) && lb_vec4_ok(dx)) {
      q.d_c = lb_make_fastdiv(c_in / 8);
      lb_launch(k_featpool_bwd_r2v8<T>, lb_grid_1d(n / 8, 256), 256, 0, lb_s(s), g, dx, (int)(n / 8), q);
    } else {
      lb_launch(k_featpool_bwd<T>, lb_grid_1d(n, 256), 256, 0, lb_s(s), g, dx, (int)n, q);
    }
  } else
    lb_launch(k_featpool_bwd_generic<T>, lb_grid_1d(n, 256), 256, 0, lb_s(s), g, dx, n, h * w, c_in, c_out, c_in / c_out);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_featpool_fwd(const void* x, void* y, int batch, int h, int w, int c_in, int c_out, int dtype, lb_stream_t s) {
  LB_REQUIRE(x && y && batch > 0 && h > 0 && w > 0 && c_out > 0 && c_in % c_out == 0);
  LB_DISPATCH(dtype, T, return featpool_fwd_t(lb_cp<T>(x), lb_p<T>(y), batch, h, w, c_in, c_out, s));
}
extern "C" int lb_featpool_bwd(const void* g, void* dx, int batch, int h, int w, int c_in, int c_out, int dtype, lb_stream_t s) {
  LB_REQUIRE(g && dx && batch > 0 && h > 0 && w > 0 && c_out > 0 && c_in % c_out == 0);
  LB_DISPATCH(dtype, T, return featpool_bwd_t(lb_cp<T>(g), lb_p<T>(dx), batch, h, w, c_in, c_out, s));
}

// ---- bilinear x2, align_corners=False (scale.py:37-38) --------------------------------------
// src = max(0, (dst + 0.5)/2 - 0.5); i0 = floor(src); i1 = min(i0+1, n-1); lambda = src - i0
__device__ __forceinline__ void lb_up2_src(int d, int n, int& i0, int& i1, float& lam) {
  float src = fmaxf(0.0f, (d + 0.5f) * 0.5f - 0.5f);
  i0 = (int)src;
  i1 = min(i0 + 1, n - 1);
  lam = src - i0;
}
// item index -> (channel group, x, y, batch) by 32-bit multiply-shift division
struct PixIdx { LbFastDiv d_c, d_w, d_h; };
__device__ __forceinline__ void pix_decode(const PixIdx& q, int i, int& ch, int& x, int& y, int& b) {
  int t;
  lb_fast_divmod(q.d_c, i, t, ch);
  lb_fast_divmod(q.d_w, t, t, x);
  lb_fast_divmod(q.d_h, t, b, y);
}
static inline PixIdx make_pix_idx(int c, int w, int h) {
  PixIdx q; q.d_c = lb_make_fastdiv(c); q.d_w = lb_make_fastdiv(w); q.d_h = lb_make_fastdiv(h);
  return q;
}
#define LB_REQUIRE_INT_ITEMS(n) LB_REQUIRE((n) < ((size_t)1 << 31) - ((size_t)1 << 24))
// V = 4: one thread = 4 consecutive channels (one 4-element access, 4x fewer index computations); V = 1: any layout.
// `c` below is the channel count in ELEMENTS; the item decode runs over channel groups of V.
template <int V> struct Acc {                 // V = LbV<T>::N: one 16-byte access of storage T
  float v[V];
  __device__ Acc() {
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = 0.0f;
  }
  template <typename T> __device__ void fma(float w, const T* p) {
    float a[V];
    lb_ldv(p, a);
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = fmaf(w, a[k], v[k]);
  }
  template <typename T> __device__ void store(T* p) const { lb_stv(p, v); }
};
template <> struct Acc<1> {
  float v;
  __device__ Acc() : v(0.0f) {}
  template <typename T> __device__ void fma(float w, const T* p) { v = fmaf(w, lb_ld1(p), v); }
  template <typename T> __device__ void store(T* p) const { lb_st1(p, v); }
};

template <typename T, int V>
__global__ void __launch_bounds__(256) k_up2_fwd(const T* __restrict__ x, T* __restrict__ y, int n_out, int h, int w, int c, const PixIdx q) {
  lb_pdl_enter();
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) {
    int ch, ox, oy, b;
    pix_decode(q, i, ch, ox, oy, b);              // q over (c / V, 2w, 2h)
    int y0, y1, x0, x1; float ly, lx;
    lb_up2_src(oy, h, y0, y1, ly);
    lb_up2_src(ox, w, x0, x1, lx);
    const T* xb = x + (size_t)b * h * w * c + ch * V;
    Acc<V> acc;
    acc.fma((1.0f - ly) * (1.0f - lx), xb + ((size_t)y0 * w + x0) * c);
    acc.fma((1.0f - ly) * lx, xb + ((size_t)y0 * w + x1) * c);
    acc.fma(ly * (1.0f - lx), xb + ((size_t)y1 * w + x0) * c);
    acc.fma(ly * lx, xb + ((size_t)y1 * w + x1) * c);
    acc.store(y + (size_t)i * V);
  }
}
// The same map, one thread per SOURCE pixel: its 2x2 output pixels need the 3x3 clamped neighbourhood (9 loads instead of
// 16, one index decode instead of four).  out[2i] = .25 x[max(i-1,0)] + .75 x[i], out[2i+1] = .75 x[i] + .25 x[min(i+1,n-1)]
// along each axis -- lb_up2_src()'s weights in closed form.  (The per-output kernel is issue-bound: ~120 instructions per
// 16-byte store, 68 % issue-slot utilisation at 2.6 TB/s.)
template <typename T, int V>
__global__ void __launch_bounds__(256) k_up2_fwd_quad(const T* __restrict__ x, T* __restrict__ y, int n_in, int h, int w, int c, const PixIdx q) {
  lb_pdl_enter();
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += stride) {
    int ch, ix, iy, b;
    pix_decode(q, i, ch, ix, iy, b);              // q over (c / V, w, h)
    const int xs[3] = {max(ix - 1, 0), ix, min(ix + 1, w - 1)};
    const int ys[3] = {max(iy - 1, 0), iy, min(iy + 1, h - 1)};
    const T* xb = x + (size_t)b * h * w * c + ch * V;
    float o[2][2][V];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const T* row = xb + (size_t)ys[r] * w * c;
      float a[V], m[V], e[V], h0[V], h1[V];
      lb_ldv(row + (size_t)xs[0] * c, a);
      lb_ldv(row + (size_t)xs[1] * c, m);
      lb_ldv(row + (size_t)xs[2] * c, e);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        h0[k] = fmaf(0.75f, m[k], 0.25f * a[k]);
        h1[k] = fmaf(0.75f, m[k], 0.25f * e[k]);
      }
#pragma unroll
      for (int k = 0; k < V; ++k) {
        if (r == 0) { o[0][0][k] = 0.25f * h0[k]; o[0][1][k] = 0.25f * h1[k]; }
        if (r == 1) {
          o[0][0][k] = fmaf(0.75f, h0[k], o[0][0][k]); o[0][1][k] = fmaf(0.75f, h1[k], o[0][1][k]);
          o[1][0][k] = 0.75f * h0[k]; o[1][1][k] = 0.75f * h1[k];
        }
        if (r == 2) { o[1][0][k] = fmaf(0.25f, h0[k], o[1][0][k]); o[1][1][k] = fmaf(0.25f, h1[k], o[1][1][k]); }
      }
    }
    T* yb = y + (((size_t)b * 2 * h + 2 * iy) * 2 * w + 2 * ix) * c + ch * V;
    lb_stv(yb, o[0][0]);
    lb_stv(yb + c, o[0][1]);
    lb_stv(yb + (size_t)2 * w * c, o[1][0]);
    lb_stv(yb + (size_t)2 * w * c + c, o[1][1]);
  }
}
// weight with which source index m receives from destination index d along one axis
__device__ __forceinline__ float lb_up2_weight(int d, int n, int m) {
  int i0, i1; float lam;
  lb_up2_src(d, n, i0, i1, lam);
  return (i0 == m ? 1.0f - lam : 0.0f) + (i1 == m ? lam : 0.0f);
}
// 1-D transposed taps of the x2 bilinear kernel: source m receives from destinations 2m-1 .. 2m+2
template <typename T, int V>
__global__ void __launch_bounds__(256) k_up2_bwd(const T* __restrict__ g, T* __restrict__ dx, int n_in, int h, int w, int c, const PixIdx q) {
  lb_pdl_enter();
  const int stride = gridDim.x * blockDim.x;
  const int ow = 2 * w, oh = 2 * h;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += stride) {
    int ch, ix, iy, b;
    pix_decode(q, i, ch, ix, iy, b);              // q over (c / V, w, h)
    // closed form of lb_up2_weight over destinations 2m-1 .. 2m+2: (.25, .75, .75, .25); at the borders the clamped tap
    // folds onto its neighbour (weight 1) and the tap outside the map is 0
    const float wy[4] = {iy > 0 ? 0.25f : 0.0f, iy > 0 ? 0.75f : 1.0f, iy < h - 1 ? 0.75f : 1.0f, iy < h - 1 ? 0.25f : 0.0f};
    const float wx[4] = {ix > 0 ? 0.25f : 0.0f, ix > 0 ? 0.75f : 1.0f, ix < w - 1 ? 0.75f : 1.0f, ix < w - 1 ? 0.25f : 0.0f};
    (void)oh; (void)ow;
    const T* gb = g + (size_t)b * oh * ow * c + ch * V;
    Acc<V> acc;
#pragma unroll
    for (int ty = 0; ty < 4; ++ty) {
      if (wy[ty] == 0.0f) continue;
      const T* row = gb + (size_t)(2 * iy - 1 + ty) * ow * c;
#pragma unroll
      for (int tx = 0; tx < 4; ++tx)
        if (wx[tx] != 0.0f) acc.fma(wy[ty] * wx[tx], row + (size_t)(2 * ix - 1 + tx) * c);
    }
    acc.store(dx + (size_t)i * V);
  }
}
template <typename T>
static int up2_fwd_t(const T* x, T* y, int batch, int h, int w, int c, lb_stream_t s) {
  const size_t n = (size_t)batch * h * w * c * 4;
  LB_REQUIRE_INT_ITEMS(n);
  constexpr int N = LbV<T>::N;
  if ((c % N) == 0 && lb_vec_ok(x) && lb_vec_ok(y))
    lb_launch(k_up2_fwd_quad<T, N>, lb_grid_1d(n / 4 / N, 256), 256, 0, lb_s(s), x, y, (int)(n / 4 / N), h, w, c, make_pix_idx(c / N, w, h));
  else
    lb_launch(k_up2_fwd<T, 1>, lb_grid_1d(n, 256), 256, 0, lb_s(s), x, y, (int)n, h, w, c, make_pix_idx(c, 2 * w, 2 * h));
  LB_LAUNCH_CHECK();
  return LB_OK;
}
template <typename T>
static int up2_bwd_t(const T* g, T* dx, int batch, int h, int w, int c, lb_stream_t s) {
  const size_t n = (size_t)batch * h * w * c;
  LB_REQUIRE_INT_ITEMS(n);
  constexpr int N = LbV<T>::N;
  if ((c % N) == 0 && lb_vec_ok(g) && lb_vec_ok(dx))
    lb_launch(k_up2_bwd<T, N>, lb_grid_1d(n / N, 256), 256, 0, lb_s(s), g, dx, (int)(n / N), h, w, c, make_pix_idx(c / N, w, h));
  else
    lb_launch(k_up2_bwd<T, 1>, lb_grid_1d(n, 256), 256, 0, lb_s(s), g, dx, (int)n, h, w, c, make_pix_idx(c, w, h));
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_upsample2x_fwd(const void* x, void* y, int batch, int h, int w, int c, int dtype, lb_stream_t s) {
  LB_REQUIRE(x && y && batch > 0 && h > 0 && w > 0 && c > 0);
  LB_DISPATCH(dtype, T, return up2_fwd_t(lb_cp<T>(x), lb_p<T>(y), batch, h, w, c, s));
}
extern "C" int lb_upsample2x_bwd(const void* g, void* dx, int batch, int h, int w, int c, int dtype, lb_stream_t s) {
  LB_REQUIRE(g && dx && batch > 0 && h > 0 && w > 0 && c > 0);
  LB_DISPATCH(dtype, T, return up2_bwd_t(lb_cp<T>(g), lb_p<T>(dx), batch, h, w, c, s));
}

// ---- AvgPool 2x2 / stride 2 (scale.py:40) ----------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256) k_avgpool2_fwd(const T* __restrict__ x, T* __restrict__ y, int n_out, int h, int w, int c, const PixIdx q) {
  lb_pdl_enter();
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) {
    int ch, ox, oy, b;
    pix_decode(q, i, ch, ox, oy, b);              // q over (c / V, w/2, h/2)
    const T* p = x + (((size_t)b * h + 2 * oy) * w + 2 * ox) * c + ch * V;
    Acc<V> acc;
    acc.fma(0.25f, p); acc.fma(0.25f, p + c); acc.fma(0.25f, p + (size_t)w * c); acc.fma(0.25f, p + (size_t)w * c + c);
    acc.store(y + (size_t)i * V);
  }
}
template <typename T, int V>
__global__ void __launch_bounds__(256) k_avgpool2_bwd(const T* __restrict__ g, T* __restrict__ dx, int n_in, int h, int w, int c, const PixIdx q) {
  lb_pdl_enter();
  const int stride = gridDim.x * blockDim.x;
  const int oh = h / 2, ow = w / 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += stride) {
    int ch, ix, iy, b;
    pix_decode(q, i, ch, ix, iy, b);              // q over (c / V, w, h)
    const int oy = iy >> 1, ox = ix >> 1;
    Acc<V> acc;
    if (oy < oh && ox < ow) acc.fma(0.25f, g + (((size_t)b * oh + oy) * ow + ox) * c + ch * V);
    acc.store(dx + (size_t)i * V);
  }
}
template <typename T>
static int avgpool2_fwd_t(const T* x, T* y, int batch, int h, int w, int c, lb_stream_t s) {
  const size_t n = (size_t)batch * (h / 2) * (w / 2) * c;
  LB_REQUIRE_INT_ITEMS(n);
  constexpr int N = LbV<T>::N;
  if ((c % N) == 0 && lb_vec_ok(x) && lb_vec_ok(y))
    lb_launch(k_avgpool2_fwd<T, N>, lb_grid_1d(n / N, 256), 256, 0, lb_s(s), x, y, (int)(n / N), h, w, c, make_pix_idx(c / N, w / 2, h / 2));
  else
    lb_launch(k_avgpool2_fwd<T, 1>, lb_grid_1d(n, 256), 256, 0, lb_s(s), x, y, (int)n, h, w, c, make_pix_idx(c, w / 2, h / 2));
  LB_LAUNCH_CHECK();
  return LB_OK;
}
template <typename T>
static int avgpool2_bwd_t(const T* g, T* dx, int batch, int h, int w, int c, lb_stream_t s) {
  const size_t n = (size_t)batch * h * w * c;
  LB_REQUIRE_INT_ITEMS(n);
  constexpr int N = LbV<T>::N;
  if ((c % N) == 0 && lb_vec_ok(g) && lb_vec_ok(dx))
    lb_launch(k_avgpool2_bwd<T, N>, lb_grid_1d(n / N, 256), 256, 0, lb_s(s), g, dx, (int)(n / N), h, w, c, make_pix_idx(c / N, w, h));
  else
    lb_launch(k_avgpool2_bwd<T, 1>, lb_grid_1d(n, 256), 256, 0, lb_s(s), g, dx, (int)n, h, w, c, make_pix_idx(c, w, h));
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_avgpool2_fwd(const void* x, void* y, int batch, int h, int w, int c, int dtype, lb_stream_t s) {
  LB_REQUIRE(x && y && batch > 0 && h > 1 && w > 1 && c > 0);
  LB_DISPATCH(dtype, T, return avgpool2_fwd_t(lb_cp<T>(x), lb_p<T>(y), batch, h, w, c, s));
}
extern "C" int lb_avgpool2_bwd(const void* g, void* dx, int batch, int h, int w, int c, int dtype, lb_stream_t s) {
  LB_REQUIRE(g && dx && batch > 0 && h > 1 && w > 1 && c > 0);
  LB_DISPATCH(dtype, T, return avgpool2_bwd_t(lb_cp<T>(g), lb_p<T>(dx), batch, h, w, c, s));
}
